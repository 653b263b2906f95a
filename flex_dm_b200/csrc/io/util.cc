// Error text, CRC-32C (hardware SSE4.2 when the CPU has it, slicing-by-8 tables otherwise), file mapping.
#include "util.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>

namespace fdio {

static thread_local char g_error[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}
const char* last_error() { return g_error; }

// CRC-32C, reflected polynomial 0x82f63b78 (iSCSI / RFC 3720; check value of "123456789" is 0xe3069283).
struct CrcTables {
  uint32_t t[8][256];
  CrcTables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1) ? 0x82f63b78u : 0u);
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
  }
};
static const CrcTables g_crc;

static uint32_t crc_sw(uint32_t crc, const uint8_t* p, size_t n) {
  uint32_t c = ~crc;
  while (n && (reinterpret_cast<uintptr_t>(p) & 7)) { c = (c >> 8) ^ g_crc.t[0][(c ^ *p++) & 0xff]; --n; }
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;
    c = g_crc.t[7][w & 0xff] ^ g_crc.t[6][(w >> 8) & 0xff] ^ g_crc.t[5][(w >> 16) & 0xff] ^ g_crc.t[4][(w >> 24) & 0xff] ^
        g_crc.t[3][(w >> 32) & 0xff] ^ g_crc.t[2][(w >> 40) & 0xff] ^ g_crc.t[1][(w >> 48) & 0xff] ^ g_crc.t[0][w >> 56];
    p += 8; n -= 8;
  }
  while (n--) c = (c >> 8) ^ g_crc.t[0][(c ^ *p++) & 0xff];
  return ~c;
}

#if defined(__x86_64__)
__attribute__((target("sse4.2"))) static uint32_t crc_hw(uint32_t crc, const uint8_t* p, size_t n) {
  uint64_t c = uint32_t(~crc);
  while (n && (reinterpret_cast<uintptr_t>(p) & 7)) { c = __builtin_ia32_crc32qi(uint32_t(c), *p++); --n; }
  while (n >= 8) { uint64_t w; memcpy(&w, p, 8); c = __builtin_ia32_crc32di(c, w); p += 8; n -= 8; }
  while (n--) c = __builtin_ia32_crc32qi(uint32_t(c), *p++);
  return ~uint32_t(c);
}
static const bool g_have_sse42 = __builtin_cpu_supports("sse4.2");
#endif

uint32_t crc32c_extend(uint32_t crc, const uint8_t* p, size_t n) {
#if defined(__x86_64__)
  if (g_have_sse42) return crc_hw(crc, p, n);
#endif
  return crc_sw(crc, p, n);
}
uint32_t crc32c_software(uint32_t crc, const uint8_t* p, size_t n) { return crc_sw(crc, p, n); }

int Mapping::open(const char* path) {
  close();
  int fd = ::open(path, O_RDONLY);
  if (fd < 0) return fail(FDIO_ERR_IO, "cannot open %s: %s", path, strerror(errno));
  struct stat st;
  if (fstat(fd, &st) != 0) { ::close(fd); return fail(FDIO_ERR_IO, "cannot stat %s: %s", path, strerror(errno)); }
  size = size_t(st.st_size);
  if (size) {
    void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) { ::close(fd); size = 0; return fail(FDIO_ERR_IO, "cannot map %s: %s", path, strerror(errno)); }
    data = static_cast<const uint8_t*>(m);
  }
  ::close(fd);
  return FDIO_OK;
}
void Mapping::close() {
  if (data) munmap(const_cast<uint8_t*>(data), size);
  data = nullptr;
  size = 0;
}

}  // namespace fdio

extern "C" {
const char* fdio_last_error(void) { return fdio::last_error(); }
int fdio_version(void) { return 1; }
uint32_t fdio_crc32c(const void* data, size_t n) { return fdio::crc32c(data, n); }
uint32_t fdio_crc32c_extend(uint32_t crc, const void* data, size_t n) { return fdio::crc32c_extend(crc, static_cast<const uint8_t*>(data), n); }
uint32_t fdio_crc32c_mask(uint32_t crc) { return fdio::crc_mask(crc); }
uint32_t fdio_crc32c_unmask(uint32_t m) { return fdio::crc_unmask(m); }
}
