#include "npz.hpp"

#include <zlib.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>

namespace viewer::npz {
namespace {

uint16_t rd16(const uint8_t *p) { return (uint16_t) (p[0] | (p[1] << 8)); }
uint32_t rd32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t) p[3] << 24); }
uint64_t rd64(const uint8_t *p) { return (uint64_t) rd32(p) | ((uint64_t) rd32(p + 4) << 32); }

[[noreturn]] void fail(const std::string &m) { throw std::runtime_error("npz: " + m); }

std::vector<uint8_t> inflate_raw(const uint8_t *src, size_t n_src, size_t n_dst) {
    std::vector<uint8_t> out(n_dst);
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -MAX_WBITS) != Z_OK) fail("inflateInit2");
    size_t done_in = 0, done_out = 0;
    int rc = Z_OK;
    while (rc != Z_STREAM_END && done_out < n_dst) {  // zlib counters are 32-bit: feed in slices
        const size_t in_now = std::min<size_t>(n_src - done_in, 1u << 30);
        const size_t out_now = std::min<size_t>(n_dst - done_out, 1u << 30);
        zs.next_in = const_cast<Bytef *>(src + done_in);
        zs.avail_in = (uInt) in_now;
        zs.next_out = out.data() + done_out;
        zs.avail_out = (uInt) out_now;
        rc = inflate(&zs, Z_NO_FLUSH);
        if (rc != Z_OK && rc != Z_STREAM_END && rc != Z_BUF_ERROR) {
            inflateEnd(&zs);
            fail("inflate failed");
        }
        done_in += in_now - zs.avail_in;
        done_out += out_now - zs.avail_out;
        if (rc == Z_BUF_ERROR && zs.avail_in == 0 && done_in >= n_src) break;
    }
    inflateEnd(&zs);
    if (done_out != n_dst) fail("inflate: short output");
    return out;
}

// CRC-32 of the member as recorded in the central directory (cnpy does not look at it; a flipped bit in a
// deflate stream can inflate "successfully").  Members above 1 GiB are checked only with MNV_NPZ_VERIFY=1:
// the check costs about as much as reading them.
void check_crc(const uint8_t *p, size_t n, uint32_t want, const std::string &name) {
    if (n > (size_t(1) << 30)) {
        const char *v = std::getenv("MNV_NPZ_VERIFY");
        if (!v || v[0] != '1') return;
    }
    uLong c = crc32(0L, Z_NULL, 0);
    while (n) {
        const size_t k = std::min<size_t>(n, 1u << 30);
        c = crc32(c, p, (uInt) k);
        p += k;
        n -= k;
    }
    if ((uint32_t) c != want) fail("CRC mismatch in member " + name);
}

}  // namespace

Array parse_npy(const uint8_t *buf, size_t len) {
    if (len < 10 || std::memcmp(buf, "\x93NUMPY", 6) != 0) fail("bad npy magic");
    const int major = buf[6];
    size_t hlen, hoff;
    if (major == 1) {
        hlen = rd16(buf + 8);
        hoff = 10;
    } else {
        if (len < 12) fail("truncated npy header");
        hlen = rd32(buf + 8);
        hoff = 12;
    }
    if (hoff + hlen > len) fail("truncated npy header");
    const std::string h(reinterpret_cast<const char *>(buf + hoff), hlen);
    Array a;
    const auto value_of = [&](const char *key) -> size_t {  // index of the first character after "key:" and blanks
        size_t p = h.find(key);
        if (p == std::string::npos) fail(std::string("npy header without ") + key);
        p = h.find(':', p);
        if (p == std::string::npos) fail(std::string("npy header: no value for ") + key);
        ++p;
        while (p < h.size() && h[p] == ' ') ++p;
        if (p >= h.size()) fail(std::string("npy header: no value for ") + key);
        return p;
    };
    // 'descr': '<f4'
    size_t p = value_of("'descr'");
    if (h[p] != '\'') fail("npy header: structured dtypes are not supported");
    const size_t q = h.find('\'', p + 1);
    if (q == std::string::npos) fail("npy header: unterminated descr");
    const std::string descr = h.substr(p + 1, q - p - 1);
    if (descr.size() < 2 || descr.size() > 16) fail("bad descr " + descr);
    size_t t = (descr[0] == '<' || descr[0] == '>' || descr[0] == '|' || descr[0] == '=') ? 1 : 0;
    if (descr[0] == '>') fail("big-endian arrays are not supported");
    a.kind = descr[t];
    if (std::string("fiub?US").find(a.kind) == std::string::npos) fail("unsupported dtype " + descr);
    size_t ws = 0;
    for (size_t i = t + 1; i < descr.size(); ++i) {
        if (!std::isdigit(static_cast<unsigned char>(descr[i]))) fail("bad descr " + descr);
        ws = ws * 10 + (size_t) (descr[i] - '0');
        if (ws > (1u << 20)) fail("bad descr " + descr);
    }
    if (a.kind == '?') ws = 1;
    const bool text = a.kind == 'U' || a.kind == 'S';
    if (ws == 0 || (!text && ws != 1 && ws != 2 && ws != 4 && ws != 8)) fail("bad item size in " + descr);
    a.word_size = a.kind == 'U' ? ws * 4 : ws;  // UCS-4 (the reference patches cnpy the same way, cnpy.cpp:113)
    // 'fortran_order': False
    a.fortran_order = h.compare(value_of("'fortran_order'"), 4, "True") == 0;
    // 'shape': (a, b, ...)
    p = value_of("'shape'");
    if (h[p] != '(') fail("npy header: shape is not a tuple");
    const size_t e = h.find(')', p);
    if (e == std::string::npos) fail("npy header: unterminated shape");
    const std::string dims = h.substr(p + 1, e - p - 1);
    size_t pos = 0, count = 1;
    while (pos < dims.size()) {
        while (pos < dims.size() && (dims[pos] == ' ' || dims[pos] == ',')) ++pos;
        if (pos >= dims.size()) break;
        size_t end = pos, v = 0;
        while (end < dims.size() && std::isdigit(static_cast<unsigned char>(dims[end]))) {
            v = v * 10 + (size_t) (dims[end] - '0');
            if (v > (size_t(1) << 48)) fail("bad shape " + dims);
            ++end;
        }
        if (end == pos || a.shape.size() >= 32) fail("bad shape " + dims);
        a.shape.push_back(v);  // 64-bit dims
        if (v != 0 && count > (size_t(1) << 56) / v) fail("bad shape " + dims);
        count *= v;
        pos = end;
    }
    if (count > (len - hoff - hlen) / a.word_size) fail("npy payload shorter than its shape");
    const size_t nbytes = count * a.word_size;
    a.bytes.assign(buf + hoff + hlen, buf + hoff + hlen + nbytes);
    return a;
}

Archive load(const std::string &path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) fail("cannot open " + path);
    const size_t size = (size_t) f.tellg();
    std::vector<uint8_t> buf(size);
    f.seekg(0);
    f.read(reinterpret_cast<char *>(buf.data()), (std::streamsize) size);
    if (!f) fail("short read " + path);
    // end of central directory (search backwards over a possible comment)
    if (size < 22) fail("not a zip file");
    size_t eocd = std::string::npos;
    for (size_t i = size - 22;; --i) {
        if (rd32(&buf[i]) == 0x06054b50) {
            eocd = i;
            break;
        }
        if (i == 0 || size - i > 22 + 65535) break;
    }
    if (eocd == std::string::npos) fail("no end-of-central-directory record");
    uint64_t n_entries = rd16(&buf[eocd + 10]), cd_off = rd32(&buf[eocd + 16]);
    if (eocd >= 20 && rd32(&buf[eocd - 20]) == 0x07064b50) {  // ZIP64 locator
        const uint64_t z64 = rd64(&buf[eocd - 20 + 8]);
        if (z64 > size || size - z64 < 56 || rd32(&buf[z64]) != 0x06064b50) fail("bad ZIP64 end record");
        n_entries = rd64(&buf[z64 + 32]);
        cd_off = rd64(&buf[z64 + 48]);
    }
    Archive out;
    size_t p = (size_t) cd_off;
    for (uint64_t i = 0; i < n_entries; ++i) {
        if (p > size || size - p < 46 || rd32(&buf[p]) != 0x02014b50) fail("bad central directory entry");
        const uint16_t method = rd16(&buf[p + 10]);
        const uint32_t want_crc = rd32(&buf[p + 16]);
        uint64_t csize = rd32(&buf[p + 20]), usize = rd32(&buf[p + 24]);
        const uint16_t nlen = rd16(&buf[p + 28]), xlen = rd16(&buf[p + 30]), clen = rd16(&buf[p + 32]);
        uint64_t lho = rd32(&buf[p + 42]);
        if (size - p - 46 < (size_t) nlen + xlen + clen) fail("central directory entry runs past the end of the file");
        std::string name(reinterpret_cast<const char *>(&buf[p + 46]), nlen);
        // ZIP64 extended information (header id 0x0001): present fields in fixed order
        size_t x = p + 46 + nlen;
        const size_t xend = x + xlen;
        while (x + 4 <= xend) {
            const uint16_t id = rd16(&buf[x]), sz = rd16(&buf[x + 2]);
            if (id == 0x0001) {
                size_t y = x + 4;
                const size_t yend = std::min(xend, y + sz);
                const auto next64 = [&](uint64_t &v) {
                    if (y + 8 > yend) fail("short ZIP64 extra field in " + name);
                    v = rd64(&buf[y]);
                    y += 8;
                };
                if (usize == 0xffffffffu) next64(usize);
                if (csize == 0xffffffffu) next64(csize);
                if (lho == 0xffffffffu) next64(lho);
            }
            x += 4 + sz;
        }
        p = xend + clen;
        if (lho > size || size - lho < 30 || rd32(&buf[lho]) != 0x04034b50) fail("bad local header for " + name);
        const size_t data = (size_t) lho + 30 + rd16(&buf[lho + 26]) + rd16(&buf[lho + 28]);
        if (data > size || csize > size - data) fail("member " + name + " runs past the end of the file");
        if (method == 0 && usize != csize) fail("stored member " + name + " with different sizes");
        if (usize > (uint64_t(1) << 40)) fail("member " + name + " is implausibly large");
        // raw deflate expands by at most ~1032x: a larger claim is a corrupt (or hostile) header, refused before
        // the output buffer is allocated
        if (method != 0 && usize / 1040 > csize + 64) fail("member " + name + " declares an impossible inflated size");
        if (name.size() > 4 && name.compare(name.size() - 4, 4, ".npy") == 0) name.resize(name.size() - 4);
        if (method == 0) {
            check_crc(&buf[data], (size_t) usize, want_crc, name);
            out[name] = parse_npy(&buf[data], (size_t) usize);
        } else if (method == 8) {
            const std::vector<uint8_t> raw = inflate_raw(&buf[data], (size_t) csize, (size_t) usize);
            check_crc(raw.data(), raw.size(), want_crc, name);
            out[name] = parse_npy(raw.data(), raw.size());
        } else {
            fail("unsupported compression method in " + name);
        }
    }
    return out;
}

namespace {
void put16(std::vector<uint8_t> &b, uint16_t v) { b.push_back(v & 0xff); b.push_back(v >> 8); }
void put32(std::vector<uint8_t> &b, uint32_t v) { for (int i = 0; i < 4; ++i) b.push_back((v >> (8 * i)) & 0xff); }
void put64(std::vector<uint8_t> &b, uint64_t v) { for (int i = 0; i < 8; ++i) b.push_back((v >> (8 * i)) & 0xff); }
uint32_t crc_of(const uint8_t *hdr, size_t nh, const void *data, size_t nd) {
    uLong c = crc32(0L, Z_NULL, 0);
    c = crc32(c, hdr, (uInt) nh);
    const uint8_t *p = static_cast<const uint8_t *>(data);
    while (nd) {  // zlib's length is 32-bit
        const size_t n = std::min<size_t>(nd, 1u << 30);
        c = crc32(c, p, (uInt) n);
        p += n;
        nd -= n;
    }
    return (uint32_t) c;
}
std::vector<uint8_t> npy_header(const Member &m) {
    std::string d = "{'descr': '" + m.descr + "', 'fortran_order': False, 'shape': (";
    for (size_t i = 0; i < m.shape.size(); ++i) d += std::to_string(m.shape[i]) + (m.shape.size() == 1 || i + 1 < m.shape.size() ? "," : "") + (i + 1 < m.shape.size() ? " " : "");
    d += "), }";
    size_t total = 10 + d.size() + 1;
    const size_t pad = (64 - total % 64) % 64;
    d.append(pad, ' ');
    d.push_back('\n');
    std::vector<uint8_t> h = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
    put16(h, (uint16_t) d.size());
    h.insert(h.end(), d.begin(), d.end());
    return h;
}
}  // namespace

void save(const std::string &path, const std::vector<Member> &members) {
    std::ofstream f(path, std::ios::binary | std::ios::trunc);
    if (!f) fail("cannot create " + path);
    struct Central { std::string name; uint32_t crc; uint64_t size, offset; };
    std::vector<Central> central;
    uint64_t offset = 0;
    for (const Member &m : members) {
        const std::vector<uint8_t> hdr = npy_header(m);
        const std::string name = m.name + ".npy";
        const uint64_t size = hdr.size() + m.nbytes;
        const uint32_t crc = crc_of(hdr.data(), hdr.size(), m.data, m.nbytes);
        std::vector<uint8_t> lh;
        put32(lh, 0x04034b50);
        put16(lh, 45);  // version needed: ZIP64
        put16(lh, 0);
        put16(lh, 0);   // stored
        put16(lh, 0);
        put16(lh, 0x21);  // DOS time / date (1980-01-01)
        put32(lh, crc);
        put32(lh, 0xffffffffu);  // sizes live in the ZIP64 extra field (like numpy's force_zip64)
        put32(lh, 0xffffffffu);
        put16(lh, (uint16_t) name.size());
        put16(lh, 20);
        lh.insert(lh.end(), name.begin(), name.end());
        put16(lh, 0x0001);
        put16(lh, 16);
        put64(lh, size);
        put64(lh, size);
        f.write(reinterpret_cast<const char *>(lh.data()), (std::streamsize) lh.size());
        f.write(reinterpret_cast<const char *>(hdr.data()), (std::streamsize) hdr.size());
        const char *p = static_cast<const char *>(m.data);
        for (size_t done = 0; done < m.nbytes;) {
            const size_t n = std::min<size_t>(m.nbytes - done, 1u << 30);
            f.write(p + done, (std::streamsize) n);
            done += n;
        }
        central.push_back({name, crc, size, offset});
        offset += lh.size() + size;
    }
    const uint64_t cd_start = offset;
    std::vector<uint8_t> cd;
    for (const Central &c : central) {
        put32(cd, 0x02014b50);
        put16(cd, 45);
        put16(cd, 45);
        put16(cd, 0);
        put16(cd, 0);
        put16(cd, 0);
        put16(cd, 0x21);
        put32(cd, c.crc);
        put32(cd, 0xffffffffu);
        put32(cd, 0xffffffffu);
        put16(cd, (uint16_t) c.name.size());
        put16(cd, 28);  // ZIP64 extra: usize, csize, offset
        put16(cd, 0);
        put16(cd, 0);
        put16(cd, 0);
        put32(cd, 0);
        put32(cd, 0xffffffffu);
        cd.insert(cd.end(), c.name.begin(), c.name.end());
        put16(cd, 0x0001);
        put16(cd, 24);
        put64(cd, c.size);
        put64(cd, c.size);
        put64(cd, c.offset);
    }
    const uint64_t cd_size = cd.size();
    // ZIP64 end of central directory + locator + classic end record
    put32(cd, 0x06064b50);
    put64(cd, 44);
    put16(cd, 45);
    put16(cd, 45);
    put32(cd, 0);
    put32(cd, 0);
    put64(cd, central.size());
    put64(cd, central.size());
    put64(cd, cd_size);
    put64(cd, cd_start);
    put32(cd, 0x07064b50);
    put32(cd, 0);
    put64(cd, cd_start + cd_size);
    put32(cd, 1);
    put32(cd, 0x06054b50);
    put16(cd, 0);
    put16(cd, 0);
    put16(cd, (uint16_t) std::min<size_t>(central.size(), 0xffff));
    put16(cd, (uint16_t) std::min<size_t>(central.size(), 0xffff));
    put32(cd, (uint32_t) std::min<uint64_t>(cd_size, 0xffffffffu));
    put32(cd, 0xffffffffu);
    put16(cd, 0);
    f.write(reinterpret_cast<const char *>(cd.data()), (std::streamsize) cd.size());
    if (!f) fail("short write " + path);
}

}  // namespace viewer::npz
