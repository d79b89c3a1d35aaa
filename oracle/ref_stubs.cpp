// oracle/ref_stubs.cpp -- TEST INFRASTRUCTURE ONLY.
// The reference's mosestokenizer.cpp builds its regexes in static initialisers that throw
// unless ../data/perluniprops is reachable from the cwd (mosestokenizer.cpp:11,78-104).
// The hot-path oracle never tokenizes text, so libbiogpt_ref.so links these two stubs
// instead of that translation unit; biogpt.cpp (gpt_tokenize / gpt_decode) only needs the
// symbols to resolve.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

std::vector<std::string> moses_tokenize(const std::string &, const std::string &) {
    fprintf(stderr, "oracle/_ref: moses_tokenize is stubbed out in the hot-path oracle build\n");
    abort();
}
std::string moses_detokenize(std::vector<std::string> &, const std::string &) {
    fprintf(stderr, "oracle/_ref: moses_detokenize is stubbed out in the hot-path oracle build\n");
    abort();
}
