// compat/mosestokenizer.h -- text pre/post-processing in front of the hot path
// (reference: /root/reference/mosestokenizer.h:17-19).  Host only.
#pragma once
#include <string>
#include <vector>

std::vector<std::string> moses_tokenize(const std::string & text, const std::string & lang);
std::string moses_detokenize(std::vector<std::string> & in_tokens, const std::string & lang);
