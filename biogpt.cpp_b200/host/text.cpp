// text.cpp -- host-side text pre/post-processing in front of the hot path: Moses-style
// tokenisation / detokenisation for English and the byte-pair merge.
//
// Reference behaviour: /root/reference/mosestokenizer.cpp (a port of sacremoses' tokenizer on
// std::regex that loads Unicode class tables from ../data at static-initialisation time) and
// /root/reference/bpe.cpp.  This is a from-scratch, table-free, single-pass implementation of
// the same English rule set on bytes: no std::regex, no data files, no static initialisers that
// can throw.  Non-ASCII bytes are treated as letters (UTF-8 words stay whole).  The three strings
// of the reference's own (never-run) unit test, mosestokenizer.cpp:491-497, are pinned in
// tests/test_host_lib.py.
#include "mosestokenizer.h"
#include "bpe.h"

#include <algorithm>
#include <cstring>
#include <set>

namespace {

inline bool is_alpha(unsigned char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c >= 0x80; }
inline bool is_digit(unsigned char c) { return c >= '0' && c <= '9'; }
inline bool is_alnum(unsigned char c) { return is_alpha(c) || is_digit(c); }
inline bool is_lower(unsigned char c) { return c >= 'a' && c <= 'z'; }
inline bool is_space(unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

// English abbreviations after which a period does not end the token (subset of Moses'
// nonbreaking_prefix.en); the second set only applies when a number follows
const std::set<std::string> & nonbreaking() {
    static const std::set<std::string> s = {
        "A","B","C","D","E","F","G","H","I","J","K","L","M","N","O","P","Q","R","S","T","U","V","W","X","Y","Z",
        "Adj","Adm","Adv","Asst","Bart","Bldg","Brig","Bros","Capt","Cmdr","Col","Comdr","Con","Corp","Cpl","DR","Dr","Drs","Ens",
        "Gen","Gov","Hon","Hr","Hosp","Insp","Lt","MM","MR","MRS","MS","Maj","Messrs","Mlle","Mme","Mr","Mrs","Ms","Msgr","Op","Ord",
        "Pfc","Ph","Prof","Pvt","Rep","Reps","Res","Rev","Rt","Sen","Sens","Sfc","Sgt","Sr","St","Supt","Surg",
        "v","vs","i.e","rev","e.g","Nos","Nr","Jan","Feb","Mar","Apr","Jun","Jul","Aug","Sep","Sept","Oct","Nov","Dec","Fig","fig","al","approx","ca","cf","etc" };
    return s;
}
const std::set<std::string> & nonbreaking_numeric() {
    static const std::set<std::string> s = { "No", "Art", "pp" };
    return s;
}

std::string collapse_spaces(const std::string & in) {
    std::string out;
    bool prev_space = true;
    for (unsigned char c : in) {
        if (c < 0x20 && !is_space(c)) continue;                  // ASCII junk
        if (is_space(c)) { if (!prev_space) out += ' '; prev_space = true; }
        else { out += (char) c; prev_space = false; }
    }
    while (!out.empty() && out.back() == ' ') out.pop_back();
    return out;
}

void replace_all(std::string & s, const std::string & from, const std::string & to) {
    for (size_t pos = 0; (pos = s.find(from, pos)) != std::string::npos; pos += to.size()) s.replace(pos, from.size(), to);
}

}  // namespace

std::vector<std::string> moses_tokenize(const std::string & text_in, const std::string & /*lang*/) {
    std::string t = collapse_spaces(text_in);

    // 1. pad every character that is neither alphanumeric nor one of  . ' ` , -  with spaces;
    //    a hyphen between two alphanumerics becomes the joiner token @-@
    std::string a;
    for (size_t i = 0; i < t.size(); i++) {
        const unsigned char c = t[i];
        if (is_alnum(c) || c == ' ' || c == '.' || c == '\'' || c == '`' || c == ',') a += (char) c;
        else if (c == '-') {
            if (i > 0 && i + 1 < t.size() && is_alnum(t[i - 1]) && is_alnum(t[i + 1])) a += " @-@ "; else a += '-';
        } else { a += ' '; a += (char) c; a += ' '; }
    }

    // 2. runs of periods ("...") stay one token
    std::string b;
    for (size_t i = 0; i < a.size(); ) {
        if (a[i] == '.' && i + 1 < a.size() && a[i + 1] == '.') {
            size_t j = i; while (j < a.size() && a[j] == '.') j++;
            b += ' '; b.append(j - i, '\x01'); b += ' ';         // \x01 stands for a protected dot
            i = j;
        } else b += a[i++];
    }

    // 3. commas: separated unless between two digits ("1,000")
    std::string c;
    for (size_t i = 0; i < b.size(); i++) {
        if (b[i] != ',') { c += b[i]; continue; }
        const bool dl = i > 0 && is_digit(b[i - 1]);
        const bool dr = i + 1 < b.size() && is_digit(b[i + 1]);
        if (dl && dr) c += ','; else c += " , ";
    }

    // 4. English apostrophes: "ain't" -> "ain 't", "'90s" / "x ' y" -> separated, "1990's" -> "1990 's"
    std::string d;
    for (size_t i = 0; i < c.size(); i++) {
        if (c[i] != '\'') { d += c[i]; continue; }
        const unsigned char l = i > 0 ? c[i - 1] : ' ', r = i + 1 < c.size() ? c[i + 1] : ' ';
        if (is_alpha(l) && is_alpha(r)) d += " '";
        else if (is_digit(l) && r == 's') d += " '";
        else d += " ' ";
    }

    // 5. word-final periods
    std::vector<std::string> words;
    { std::string w; for (char ch : d) { if (ch == ' ') { if (!w.empty()) words.push_back(w); w.clear(); } else w += ch; } if (!w.empty()) words.push_back(w); }
    std::vector<std::string> out;
    for (size_t i = 0; i < words.size(); i++) {
        std::string & w = words[i];
        if (w.size() > 1 && w.back() == '.') {
            const std::string pre = w.substr(0, w.size() - 1);
            const bool inner_dot_alpha = pre.find('.') != std::string::npos && std::any_of(pre.begin(), pre.end(), [](char ch) { return is_alpha((unsigned char) ch); });
            const bool next_lower = i + 1 < words.size() && is_lower((unsigned char) words[i + 1][0]);
            const bool next_digit = i + 1 < words.size() && is_digit((unsigned char) words[i + 1][0]);
            if (inner_dot_alpha || nonbreaking().count(pre) || next_lower || (nonbreaking_numeric().count(pre) && next_digit)) out.push_back(w);
            else { out.push_back(pre); out.push_back("."); }
        } else out.push_back(w);
    }

    // 6. restore protected dots, escape the characters Moses reserves
    for (std::string & w : out) {
        std::replace(w.begin(), w.end(), '\x01', '.');
        replace_all(w, "&", "&amp;");  replace_all(w, "|", "&#124;"); replace_all(w, "<", "&lt;");   replace_all(w, ">", "&gt;");
        replace_all(w, "'", "&apos;"); replace_all(w, "\"", "&quot;"); replace_all(w, "[", "&#91;"); replace_all(w, "]", "&#93;");
    }
    return out;
}

std::string moses_detokenize(std::vector<std::string> & in_tokens, const std::string & /*lang*/) {
    std::string text;
    bool glue_next = false;          // no space before the next token
    int dq = 0, sq = 0;              // quote parity
    for (size_t i = 0; i < in_tokens.size(); i++) {
        std::string tok = in_tokens[i];
        replace_all(tok, "&bar;", "|"); replace_all(tok, "&#124;", "|"); replace_all(tok, "&lt;", "<"); replace_all(tok, "&gt;", ">");
        replace_all(tok, "&bra;", "["); replace_all(tok, "&ket;", "]"); replace_all(tok, "&quot;", "\""); replace_all(tok, "&apos;", "'");
        replace_all(tok, "&#91;", "["); replace_all(tok, "&#93;", "]"); replace_all(tok, "&amp;", "&");
        if (tok.empty()) continue;
        if (tok == "@-@") { text += '-'; glue_next = true; continue; }
        bool attach_left = false, glue_after = false;
        const unsigned char f = tok[0];
        if (tok.size() == 1 && strchr(".,:;?!%)]}", f)) attach_left = true;
        else if (tok.size() == 1 && strchr("([{$#", f)) glue_after = true;
        else if (tok == "\"") { if (dq++ % 2 == 0) glue_after = true; else attach_left = true; }
        else if (tok == "'") { if (sq++ % 2 == 0) glue_after = true; else attach_left = true; }
        else if (f == '\'' && tok.size() > 1 && is_alpha((unsigned char) tok[1])) attach_left = true;     // 's 't 're ...
        if (!text.empty() && !glue_next && !attach_left) text += ' ';
        text += tok;
        glue_next = glue_after;
    }
    return collapse_spaces(text);
}

// ------------------------------------------------------------------------------------------------
// byte-pair merge.  Symbols start as single bytes, the last one carries "</w>"; repeatedly the
// adjacent pair with the lowest rank is merged (all its occurrences, left to right) until no
// ranked pair is left.
// ------------------------------------------------------------------------------------------------
std::string bpe(const std::string & token, std::map<word_pair, int> & bpe_ranks) {
    if (token.empty()) return "</w>";
    std::vector<std::string> sym;
    for (size_t i = 0; i + 1 < token.size(); i++) sym.emplace_back(1, token[i]);
    sym.push_back(token.substr(token.size() - 1) + "</w>");
    while (sym.size() > 1) {
        int best_rank = -1; word_pair best;
        for (size_t i = 0; i + 1 < sym.size(); i++) {
            auto it = bpe_ranks.find(word_pair(sym[i], sym[i + 1]));
            if (it != bpe_ranks.end() && (best_rank < 0 || it->second < best_rank)) { best_rank = it->second; best = it->first; }
        }
        if (best_rank < 0) break;
        std::vector<std::string> merged;
        for (size_t i = 0; i < sym.size(); ) {
            if (i + 1 < sym.size() && sym[i] == best.first && sym[i + 1] == best.second) { merged.push_back(sym[i] + sym[i + 1]); i += 2; }
            else merged.push_back(sym[i++]);
        }
        sym.swap(merged);
    }
    std::string out;
    for (size_t i = 0; i < sym.size(); i++) { if (i) out += ' '; out += sym[i]; }
    if (out == "\n  </w>") out = "\n</w>";
    return out;
}
