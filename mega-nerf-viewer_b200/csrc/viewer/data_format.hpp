// viewer::DataFormat — same public surface as the reference's include/data_format.hpp:7-22.
#pragma once
#include <string>

namespace viewer {

struct DataFormat {
    enum { RGBA, SH, _COUNT } format = RGBA;
    int basis_dim = -1;  // SH dimension per channel, -1 when absent

    void parse(const std::string &str);  // "SH9" -> {SH, 9}; "RGBA" -> {RGBA, -1}
    std::string to_string() const;
};

}  // namespace viewer
