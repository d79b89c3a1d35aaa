// text.cpp -- host-side text pre/post-processing in front of the hot path: Moses tokenisation / detokenisation and the
// byte-pair merge.
//
// Reference behaviour: /root/reference/mosestokenizer.cpp (sacremoses' tokenizer re-stated with std::regex over std::string)
// and /root/reference/bpe.cpp.  What that code computes AS BUILT is byte-level: the character classes are bracket expressions
// made of the raw bytes of data/perluniprops/Is{Alnum,Alpha,Lower,N,Sc}.txt, so a class is a set of BYTES (every ASCII member
// plus every byte that occurs in the UTF-8 encoding of a member), and each rule is a short sequence of such classes.  This file
// is a from-scratch scanner for the same rule sequence on the same byte sets -- no std::regex (the reference spends seconds
// constructing 30 KB bracket expressions in static initialisers that throw when ../data is not reachable), no exceptions:
//   * the five byte sets come from ../data/perluniprops/*.txt when that directory is reachable from the cwd (as the reference
//     requires), else from the built-in masks below (the same sets, derived from those files);
//   * the non-breaking prefixes come from ../data/nonbreaking_prefixes/nonbreaking_prefix.<lang> when reachable, else from the
//     built-in English list.  As built, the reference strips everything after the first '#' of a line, so the #NUMERIC_ONLY#
//     marker never survives and "No", "Art", "pp" are ordinary prefixes; its "next token starts lower-case" exception is dead
//     code (mosestokenizer.cpp:259 builds a string of N copies of byte 1); runs of dots are only partly protected
//     (mosestokenizer.cpp:181-199).  All of that is reproduced: tests/test_text.py compares token lists with the reference's
//     own output (tests/golden/moses_golden.json, and live against the reference where /root/reference exists).
#include "mosestokenizer.h"
#include "bpe.h"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <sstream>

namespace {

struct ByteSet {
    uint32_t w[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    bool has(unsigned char c) const { return (w[c >> 5] >> (c & 31)) & 1u; }
    void add(unsigned char c) { w[c >> 5] |= 1u << (c & 31); }
    void add(const char * chars) { for (const char * p = chars; *p; p++) add((unsigned char) *p); }
    void add_range(int lo, int hi) { for (int c = lo; c <= hi; c++) add((unsigned char) c); }
    void remove(std::initializer_list<int> cs) { for (int c : cs) w[c >> 5] &= ~(1u << (c & 31)); }
};

struct Tables {
    ByteSet alnum, alpha, lower, num, sc;
    Tables() {
        // built-in: the byte sets of the reference's data files (ASCII members + all UTF-8 bytes that occur in them)
        alpha.add_range('A', 'Z'); alpha.add_range('a', 'z'); alpha.add_range(0x80, 0xEF);
        alpha.remove({ 0xC0, 0xC1, 0xCC, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEB, 0xEC, 0xEE });
        alnum = alpha; alnum.add_range('0', '9');
        lower.add_range('a', 'z'); lower.add_range(0x80, 0xF0);
        lower.remove({ 0xC0, 0xC1, 0xCC, 0xD7, 0xD8, 0xD9, 0xDA, 0xDB, 0xDC, 0xDD, 0xDE, 0xDF, 0xE0, 0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEB, 0xEC, 0xED, 0xEE });
        num.add_range('0', '9'); num.add('\n'); num.add_range(0x80, 0xEF);
        num.remove({ 0xC0, 0xC1, 0xC3, 0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xCB, 0xCC, 0xCD, 0xCE, 0xCF, 0xD0, 0xD1, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8,
                     0xDA, 0xDC, 0xDD, 0xDE, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEB, 0xEC, 0xED, 0xEE });
        sc.add('$'); sc.add('\n');
        for (int c : { 0x82, 0x84, 0x8B, 0x8F, 0x9B, 0x9F, 0xA0, 0xA1, 0xA2, 0xA3, 0xA4, 0xA5, 0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xAB, 0xAC, 0xAD, 0xAE, 0xAF,
                       0xB0, 0xB1, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA, 0xBB, 0xBC, 0xBD, 0xBF, 0xC2, 0xD6, 0xD8, 0xE0, 0xE1, 0xE2, 0xEA, 0xEF }) sc.add((unsigned char) c);
        // the reference's own source of truth, when it is where the reference looks for it
        load("IsAlnum", alnum); load("IsAlpha", alpha); load("IsLower", lower); load("IsN", num); load("IsSc", sc);
    }
    static void load(const char * category, ByteSet & set) {
        std::ifstream in(std::string("../data/perluniprops/") + category + ".txt", std::ios::binary);
        if (!in) return;
        ByteSet s;
        bool any = false;
        for (std::istreambuf_iterator<char> it(in), end; it != end; ++it) { s.add((unsigned char) *it); any = true; }
        if (any) set = s;
    }
};
const Tables & tables() { static const Tables t; return t; }

inline bool is_space(unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

// nonbreaking_prefix.<lang>, as the reference reads it: everything after the first '#' of a line is dropped, then trimmed
std::vector<std::string> nonbreaking_prefixes(const std::string & lang) {
    std::vector<std::string> out;
    std::ifstream in("../data/nonbreaking_prefixes/nonbreaking_prefix." + (lang.empty() ? std::string("en") : lang));
    if (in) {
        for (std::string line; std::getline(in, line); ) {
            line = line.substr(0, line.find('#'));
            size_t a = 0, b = line.size();
            while (a < b && is_space((unsigned char) line[a])) a++;
            while (b > a && is_space((unsigned char) line[b - 1])) b--;
            if (b > a) out.push_back(line.substr(a, b - a));
        }
        return out;
    }
    if (lang != "en" && !lang.empty()) return out;                // no table for that language: every final period is split off
    static const char * const en[] = {
        "A","B","C","D","E","F","G","H","I","J","K","L","M","N","O","P","Q","R","S","T","U","V","W","X","Y","Z",
        "Adj","Adm","Adv","Asst","Bart","Bldg","Brig","Bros","Capt","Cmdr","Col","Comdr","Con","Corp","Cpl","DR","Dr","Drs","Ens",
        "Gen","Gov","Hon","Hr","Hosp","Insp","Lt","MM","MR","MRS","MS","Maj","Messrs","Mlle","Mme","Mr","Mrs","Ms","Msgr","Op","Ord",
        "Pfc","Ph","Prof","Pvt","Rep","Reps","Res","Rev","Rt","Sen","Sens","Sfc","Sgt","Sr","St","Supt","Surg",
        "v","vs","i.e","rev","e.g","Rs","No","Nos","Art","Nr","pp","Jan","Feb","Mar","Apr","Jun","Jul","Aug","Sep","Oct","Nov","Dec" };
    out.assign(std::begin(en), std::end(en));
    return out;
}

// runs of white space -> one space; leading / trailing white space removed
std::string squeeze(const std::string & in) {
    std::string out;
    bool pending = false;
    for (unsigned char c : in) {
        if (is_space(c)) { pending = !out.empty(); continue; }
        if (pending) { out += ' '; pending = false; }
        out += (char) c;
    }
    return out;
}

void replace_all(std::string & s, const std::string & from, const std::string & to) {
    for (size_t pos = 0; (pos = s.find(from, pos)) != std::string::npos; pos += to.size()) s.replace(pos, from.size(), to);
}

std::vector<std::string> split_ws(const std::string & s) {
    std::vector<std::string> v; std::string w;
    for (unsigned char c : s) { if (is_space(c)) { if (!w.empty()) v.push_back(w); w.clear(); } else w += (char) c; }
    if (!w.empty()) v.push_back(w);
    return v;
}

// One left-to-right, non-overlapping pass of a three-byte rule  X ' Y -> X <mid> Y  (the apostrophe rules): `l` / `r` decide on
// the bytes around the apostrophe.
template <typename L, typename R>
std::string apostrophe_pass(const std::string & s, L l, R r, const char * mid) {
    std::string out;
    const size_t n = s.size();
    for (size_t i = 0; i < n; ) {
        if (i + 2 < n && s[i + 1] == '\'' && l((unsigned char) s[i]) && r((unsigned char) s[i + 2])) {
            out += s[i]; out += mid; out += s[i + 2];
            i += 3;
        } else out += s[i++];
    }
    return out;
}

}  // namespace

// test door (host_capi.cpp): the 256-bit membership mask of a byte class, 0 IsAlnum 1 IsAlpha 2 IsLower 3 IsN 4 IsSc
void bgpt_text_class_mask(int which, uint32_t out[8]) {
    const Tables & T = tables();
    const ByteSet * s[5] = { &T.alnum, &T.alpha, &T.lower, &T.num, &T.sc };
    for (int i = 0; i < 8; i++) out[i] = (which >= 0 && which < 5) ? s[which]->w[i] : 0u;
}

std::vector<std::string> moses_tokenize(const std::string & text_in, const std::string & lang) {
    const Tables & T = tables();
    // white space runs -> one space, ASCII control bytes dropped, ends trimmed            (mosestokenizer.cpp:303-309)
    std::string t;
    {
        std::string a;
        bool in_ws = false;
        for (unsigned char c : text_in) { if (is_space(c)) { if (!in_ws) a += ' '; in_ws = true; } else { a += (char) c; in_ws = false; } }
        for (unsigned char c : a) if (c > 0x1F) t += (char) c;
        size_t b = 0, e = t.size();
        while (b < e && is_space((unsigned char) t[b])) b++;
        while (e > b && is_space((unsigned char) t[e - 1])) e--;
        t = t.substr(b, e - b);
    }
    // every byte outside IsAlnum, white space and  . ' ` , -  is padded with spaces                          (:312)
    {
        std::string a;
        for (unsigned char c : t) {
            if (T.alnum.has(c) || is_space(c) || c == '.' || c == '\'' || c == '`' || c == ',' || c == '-') a += (char) c;
            else { a += ' '; a += (char) c; a += ' '; }
        }
        t.swap(a);
    }
    // a hyphen between two alphanumeric bytes becomes the joiner token @-@ (the right neighbour is only looked at)   (:315)
    {
        std::string a;
        const size_t n = t.size();
        for (size_t i = 0; i < n; ) {
            if (i + 2 < n && t[i + 1] == '-' && T.alnum.has((unsigned char) t[i]) && T.alnum.has((unsigned char) t[i + 2])) { a += t[i]; a += " @-@ "; i += 2; }
            else a += t[i++];
        }
        t.swap(a);
    }
    // runs of dots: ".." + more -> DOTMULTI + the rest of the run; then "DOTMULTI." followed by a non-dot byte -> "DOTDOTMULTI " +
    // that byte; then every remaining "DOTMULTI." -> "DOTDOTMULTI" (one pass each, as built)                 (:181-199)
    {
        std::string a;
        const size_t n = t.size();
        for (size_t i = 0; i < n; ) {
            if (t[i] == '.' && i + 1 < n && t[i + 1] == '.') {
                size_t j = i + 1; while (j < n && t[j] == '.') j++;
                a += "DOTMULTI"; a.append(j - i - 1, '.');
                i = j;
            } else a += t[i++];
        }
        std::string b;
        const size_t m = a.size();
        for (size_t i = 0; i < m; ) {
            if (a.compare(i, 9, "DOTMULTI.") == 0 && i + 9 < m && a[i + 9] != '.') { b += "DOTDOTMULTI "; b += a[i + 9]; i += 10; }
            else b += a[i++];
        }
        replace_all(b, "DOTMULTI.", "DOTDOTMULTI");
        t.swap(b);
    }
    // commas are separated unless a numeric byte stands on that side                                          (:321-326)
    {
        std::string a;
        size_t n = t.size();
        for (size_t i = 0; i < n; ) {
            if (i + 1 < n && t[i + 1] == ',' && !T.num.has((unsigned char) t[i])) { a += t[i]; a += " , "; i += 2; }
            else a += t[i++];
        }
        std::string b;
        n = a.size();
        for (size_t i = 0; i < n; ) {
            if (a[i] == ',' && i + 1 < n && !T.num.has((unsigned char) a[i + 1])) { b += " , "; b += a[i + 1]; i += 2; }
            else b += a[i++];
        }
        n = b.size();
        if (n >= 2 && b[n - 1] == ',' && T.num.has((unsigned char) b[n - 2])) { b.erase(n - 1); b += " , "; }
        t.swap(b);
    }
    // apostrophes                                                                                              (:329-343)
    if (lang == "en") {
        auto al = [&](unsigned char c) { return T.alpha.has(c); };
        auto nal = [&](unsigned char c) { return !T.alpha.has(c); };
        auto naln = [&](unsigned char c) { return !T.alpha.has(c) && !T.num.has(c); };
        auto nu = [&](unsigned char c) { return T.num.has(c); };
        auto ess = [](unsigned char c) { return c == 's'; };
        t = apostrophe_pass(t, nal, nal, " ' ");
        t = apostrophe_pass(t, naln, al, " ' ");
        t = apostrophe_pass(t, al, nal, " ' ");
        t = apostrophe_pass(t, al, al, " '");
        t = apostrophe_pass(t, nu, ess, " '");
    } else if (lang == "fr") {
        auto al = [&](unsigned char c) { return T.alpha.has(c); };
        auto nal = [&](unsigned char c) { return !T.alpha.has(c); };
        t = apostrophe_pass(t, nal, nal, " ' ");
        t = apostrophe_pass(t, nal, al, " ' ");
        t = apostrophe_pass(t, al, nal, " ' ");
        t = apostrophe_pass(t, al, al, "' ");
    } else replace_all(t, "'", " ' ");
    // word-final periods: kept on a token whose stem has an inner period and a letter, or is a non-breaking prefix   (:238-290)
    {
        const std::vector<std::string> nb = nonbreaking_prefixes(lang);
        std::vector<std::string> words = split_ws(t);
        std::string a;
        for (std::string & w : words) {
            if (w.size() > 1 && w.back() == '.') {
                const std::string pre = w.substr(0, w.size() - 1);
                const bool inner = pre.find('.') != std::string::npos &&
                                   std::any_of(pre.begin(), pre.end(), [&](char ch) { return T.alpha.has((unsigned char) ch); });
                if (!inner && std::find(nb.begin(), nb.end(), pre) == nb.end()) w = pre + " .";
            }
            if (!a.empty()) a += ' ';
            a += w;
        }
        t.swap(a);
    }
    t = squeeze(t);
    // a final  .'  is split                                                                                    (:352)
    if (t.size() >= 2 && t.compare(t.size() - 2, 2, ".'") == 0) { t.erase(t.size() - 2); t += " . ' "; }
    // the protected dots come back, then the characters Moses reserves are escaped                            (:201-216, 355-358)
    while (t.find("DOTDOTMULTI") != std::string::npos) replace_all(t, "DOTDOTMULTI", "DOTMULTI.");
    replace_all(t, "DOTMULTI", ".");
    replace_all(t, "&", "&amp;");  replace_all(t, "|", "&#124;"); replace_all(t, "<", "&lt;");   replace_all(t, ">", "&gt;");
    replace_all(t, "'", "&apos;"); replace_all(t, "\"", "&quot;"); replace_all(t, "[", "&#91;"); replace_all(t, "]", "&#93;");
    return split_ws(t);
}

// as built (mosestokenizer.cpp:370-466): the XML entities are NOT unescaped (the reference discards the result of its
// regex_replace), the "closing punctuation" rule only matches the literal token "[,.?!:;\%}]" followed by ')' characters, and
// " @-@" (with its leading space only) becomes "-".
std::string moses_detokenize(std::vector<std::string> & in_tokens, const std::string & lang) {
    const Tables & T = tables();
    std::string text = " ";
    for (const std::string & tok : in_tokens) { text += tok; text += ' '; }
    replace_all(text, " @-@", "-");
    const std::vector<std::string> tokens = split_ws(text);
    auto all_in = [](const std::string & s, const ByteSet & set) { return !s.empty() && std::all_of(s.begin(), s.end(), [&](char c) { return set.has((unsigned char) c); }); };
    ByteSet open = T.sc; open.add("([{"); open.add((unsigned char) 0xC2); open.add((unsigned char) 0xBF); open.add((unsigned char) 0xA1);
    ByteSet quotes; quotes.add("'\"`"); for (int c : { 0xE2, 0x80, 0x9E, 0x9C }) quotes.add((unsigned char) c);
    ByteSet fancy; for (int c : { 0xE2, 0x80, 0x9E, 0x9C, 0x9D }) fancy.add((unsigned char) c);
    auto literal_then = [](const std::string & s, const char * lit, char rep) {          // s == lit + rep{1,}
        const size_t n = strlen(lit);
        if (s.size() <= n || s.compare(0, n, lit) != 0) return false;
        return std::all_of(s.begin() + n, s.end(), [&](char c) { return c == rep; });
    };
    std::vector<std::pair<std::string, int>> quote_counts = { { "'", 0 }, { "\"", 0 }, { "``", 0 }, { "`", 0 }, { "''", 0 } };
    auto count_of = [&](const std::string & q) -> int & {
        for (auto & kv : quote_counts) if (kv.first == q) return kv.second;
        quote_counts.emplace_back(q, 0);
        return quote_counts.back().second;
    };
    std::string prepend = " ", out;
    for (size_t i = 0; i < tokens.size(); i++) {
        const std::string & tok = tokens[i];
        if (all_in(tok, open)) { out += prepend + tok; prepend = ""; }
        else if (literal_then(tok, "[,.?!:;\\%}]", ')')) {
            if (lang == "fr" && literal_then(tok, "[?!:;\\%", ']')) out += ' ';
            out += tok; prepend = " ";
        }
        else if (lang == "en" && i > 0 && tok.size() >= 2 && tok[0] == '\'' && T.alpha.has((unsigned char) tok[1])) { out += tok; prepend = " "; }
        else if (lang == "fr" || lang == "it" || lang == "ga") {
            if (i + 2 <= tokens.size() - 0 && i + 1 < tokens.size() && tok.size() >= 2 && tok.back() == '\'' && T.alpha.has((unsigned char) tok[tok.size() - 2]) &&
                T.alpha.has((unsigned char) tokens[i + 1][0])) { out += prepend + tok; prepend = ""; }
        }
        else if (all_in(tok, quotes)) {
            const std::string norm = all_in(tok, fancy) ? std::string("\"") : tok;
            int & cnt = count_of(norm);
            if (cnt % 2 == 0) {
                if (lang == "en" && tok == "'" && i > 0 && !tokens[i - 1].empty() && tokens[i - 1].back() == 's') { out += tok; prepend = " "; }
                else { out += prepend + tok; prepend = ""; cnt += 1; }
            } else { out += tok; prepend = " "; cnt += 1; }
        }
        else { out += prepend + tok; prepend = " "; }
    }
    // two or more spaces -> one; trimmed
    std::string res;
    for (size_t i = 0; i < out.size(); ) {
        if (out[i] == ' ') { size_t j = i; while (j < out.size() && out[j] == ' ') j++; res += ' '; i = j; }
        else res += out[i++];
    }
    size_t b = 0, e = res.size();
    while (b < e && is_space((unsigned char) res[b])) b++;
    while (e > b && is_space((unsigned char) res[e - 1])) e--;
    return res.substr(b, e - b);
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
