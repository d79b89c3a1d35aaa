// oracle/ref_moses_shim.cpp -- TEST INFRASTRUCTURE: extern "C" doors onto the UNMODIFIED reference tokenizer
// (/root/reference/mosestokenizer.cpp), compiled by oracle/Makefile into _ref/libmoses_ref.so.  The reference builds its regexes in
// static initialisers that read ../data/perluniprops relative to the cwd (mosestokenizer.cpp:11, 78-104) and re-reads
// ../data/nonbreaking_prefixes on every call, so the library must be loaded AND called with the cwd inside a sub-directory of the
// reference tree (tests/golden/make_moses_golden.py, tests/test_text.py do that).
#include "mosestokenizer.h"
#include <cstring>

extern "C" {
int refmoses_tokenize(const char * text, const char * lang, char * out, int cap) {
    // as built the reference THROWS (std::length_error, mosestokenizer.cpp:259: std::string(count = a negative char, 1)) when a
    // token that ends in a period is followed by a token whose first byte is >= 0x80: reported as -2
    std::vector<std::string> v;
    try { v = moses_tokenize(text, lang); } catch (const std::exception &) { return -2; }
    std::string s;
    for (size_t i = 0; i < v.size(); i++) { if (i) s += '\n'; s += v[i]; }
    if ((int) s.size() + 1 > cap) return -1;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int) v.size();
}
int refmoses_detokenize(const char * tokens_nl, const char * lang, char * out, int cap) {
    std::vector<std::string> v; std::string cur;
    for (const char * p = tokens_nl; *p; p++) { if (*p == '\n') { v.push_back(cur); cur.clear(); } else cur += *p; }
    if (!cur.empty() || !v.empty()) v.push_back(cur);
    const std::string s = moses_detokenize(v, lang);
    if ((int) s.size() + 1 > cap) return -1;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int) s.size();
}
}
