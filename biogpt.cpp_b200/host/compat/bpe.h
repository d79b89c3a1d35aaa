// compat/bpe.h -- byte-pair merge of one Moses token (reference: /root/reference/bpe.h:9)
#pragma once
#include <map>
#include <string>

typedef std::pair<std::string, std::string> word_pair;

// returns the sub-words separated by single spaces, the last one carrying "</w>"
std::string bpe(const std::string & token, std::map<word_pair, int> & bpe_ranks);
