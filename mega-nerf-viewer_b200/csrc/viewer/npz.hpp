// Minimal .npy/.npz reader for the keys N3Tree::load_npz reads (the reference vendors a
// patched cnpy, 3rdparty/cnpy/cnpy.cpp:303-369).  Entries are located through the ZIP
// central directory (so data descriptors and ZIP64 sizes are handled uniformly), stored
// and raw-deflate members are supported, npy format versions 1-3.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace viewer::npz {

struct Array {
    std::vector<size_t> shape;
    size_t word_size = 0;  // bytes per element ('U' = 4 per character)
    char kind = 0;         // numpy type char: f, i, u, b, U, ...
    bool fortran_order = false;
    std::vector<uint8_t> bytes;

    size_t num_vals() const {
        size_t n = 1;
        for (size_t s : shape) n *= s;
        return n;
    }
    template <typename T>
    const T *data() const { return reinterpret_cast<const T *>(bytes.data()); }
    // Size- and width-checked view: at least n elements of sizeof(T) bytes each, else std::runtime_error.
    template <typename T>
    const T *checked(size_t n, const char *what) const {
        if (word_size != sizeof(T) || n > bytes.size() / sizeof(T) || num_vals() < n)
            throw std::runtime_error(std::string("npz: ") + what + ": expected at least " + std::to_string(n) + " elements of " +
                                     std::to_string(sizeof(T)) + " bytes, found " + std::to_string(num_vals()) + " of " +
                                     std::to_string(word_size));
        return data<T>();
    }
};

using Archive = std::map<std::string, Array>;

// Throws std::runtime_error on malformed input.
Archive load(const std::string &path);

// Writer side (the reference only reads): `.npy` v1 members, stored (no compression), ZIP64 records
// whenever a size or offset needs them, so multi-GB refined trees can be written.
struct Member {
    std::string name;            // without ".npy"
    std::string descr;           // numpy dtype string, e.g. "<f2", "<i4", "<U3"
    std::vector<size_t> shape;   // {} = scalar
    const void *data;
    size_t nbytes;
};
void save(const std::string &path, const std::vector<Member> &members);
Array parse_npy(const uint8_t *buf, size_t len);

}  // namespace viewer::npz
