// Shared helpers of libflexdm_io: thread-local error text, CRC-32C, protobuf wire reader/writer, mmap.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <string_view>

#include "../../../include/flexdm_io.h"

namespace fdio {

int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
const char* last_error();

uint32_t crc32c_extend(uint32_t crc, const uint8_t* p, size_t n);
inline uint32_t crc32c(const void* p, size_t n) { return crc32c_extend(0, static_cast<const uint8_t*>(p), n); }
inline uint32_t crc_mask(uint32_t c) { return ((c >> 15) | (c << 17)) + 0xa282ead8u; }
inline uint32_t crc_unmask(uint32_t m) { uint32_t r = m - 0xa282ead8u; return (r >> 17) | (r << 15); }

// Read-only file mapping (TFRecord shards, checkpoint shards).
struct Mapping {
  const uint8_t* data = nullptr;
  size_t size = 0;
  int open(const char* path);
  void close();
  ~Mapping() { close(); }
  Mapping() = default;
  Mapping(const Mapping&) = delete;
  Mapping& operator=(const Mapping&) = delete;
};

// Protocol-buffers wire format, just what the two formats need: varints, tags, length-delimited fields.
struct Wire {
  const uint8_t* p;
  const uint8_t* end;
  Wire(const uint8_t* b, size_t n) : p(b), end(b + n) {}
  bool done() const { return p >= end; }
  bool varint(uint64_t* v) {
    uint64_t r = 0;
    for (int shift = 0; shift < 64 && p < end; shift += 7) {
      uint8_t b = *p++;
      r |= uint64_t(b & 0x7f) << shift;
      if (!(b & 0x80)) { *v = r; return true; }
    }
    return false;
  }
  bool tag(uint32_t* field, uint32_t* type) {
    uint64_t t;
    if (!varint(&t)) return false;
    *field = uint32_t(t >> 3);
    *type = uint32_t(t & 7);
    return true;
  }
  bool bytes(const uint8_t** b, size_t* n) {
    uint64_t len;
    if (!varint(&len) || len > size_t(end - p)) return false;
    *b = p; *n = size_t(len); p += len;
    return true;
  }
  bool fixed32(uint32_t* v) { if (end - p < 4) return false; memcpy(v, p, 4); p += 4; return true; }
  bool fixed64(uint64_t* v) { if (end - p < 8) return false; memcpy(v, p, 8); p += 8; return true; }
  bool skip(uint32_t type) {
    uint64_t v; uint32_t w; const uint8_t* b; size_t n;
    switch (type) {
      case 0: return varint(&v);
      case 1: return fixed64(&v);
      case 2: return bytes(&b, &n);
      case 5: return fixed32(&w);
      default: return false;  // groups are not used by either format
    }
  }
};

inline void put_varint(std::string* s, uint64_t v) {
  while (v >= 0x80) { s->push_back(char(v | 0x80)); v >>= 7; }
  s->push_back(char(v));
}
inline void put_tag(std::string* s, uint32_t field, uint32_t type) { put_varint(s, (uint64_t(field) << 3) | type); }
inline void put_bytes(std::string* s, uint32_t field, std::string_view b) { put_tag(s, field, 2); put_varint(s, b.size()); s->append(b); }
inline void put_fixed32(std::string* s, uint32_t v) { s->append(reinterpret_cast<const char*>(&v), 4); }
inline void put_fixed64(std::string* s, uint64_t v) { s->append(reinterpret_cast<const char*>(&v), 8); }

}  // namespace fdio
