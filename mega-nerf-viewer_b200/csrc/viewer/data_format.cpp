#include "data_format.hpp"

#include <cctype>
#include <cstdlib>

namespace viewer {

// Behaviour of src/data_format.cpp:5-24: the alphabetic prefix names the format ("SH",
// anything else is RGBA), the remainder is the basis dimension.
void DataFormat::parse(const std::string &str) {
    size_t split = 0;
    while (split < str.size() && std::isalpha(static_cast<unsigned char>(str[split]))) ++split;
    if (split == str.size()) {
        format = RGBA;
        basis_dim = -1;
        return;
    }
    basis_dim = std::atoi(str.c_str() + split);
    format = str.compare(0, split, "SH") == 0 ? SH : RGBA;
}

std::string DataFormat::to_string() const {
    std::string name = format == SH ? "SH" : (format == RGBA ? "RGBA" : "UNKNOWN");
    if (basis_dim != -1) name += std::to_string(basis_dim);
    return name;
}

}  // namespace viewer
