// TFRecord framing: what tf.data.TFRecordDataset reads for DataSpec.make_dataset (src/mfp/mfp/data/spec.py:231-236).
// Published format (tensorflow/core/lib/io/record_writer.h): per record
//   uint64 length | uint32 masked_crc32c(length) | byte data[length] | uint32 masked_crc32c(data)     (little endian)
#include <cerrno>
#include <vector>

#include "util.hpp"

struct fdio_tfrecord {
  fdio::Mapping map;
  std::vector<uint64_t> offset;  // payload offsets
  std::vector<uint64_t> length;
};

extern "C" {

fdio_tfrecord* fdio_tfrecord_open(const char* path, int verify_crc) {
  if (!path) { fdio::fail(FDIO_ERR_ARG, "fdio_tfrecord_open: null path"); return nullptr; }
  auto* f = new fdio_tfrecord;
  if (f->map.open(path) != FDIO_OK) { delete f; return nullptr; }
  const uint8_t* d = f->map.data;
  const size_t n = f->map.size;
  size_t pos = 0;
  while (pos < n) {
    if (n - pos < 12) { fdio::fail(FDIO_ERR_CORRUPT, "%s: truncated record header at byte %zu", path, pos); delete f; return nullptr; }
    uint64_t len;
    uint32_t len_crc;
    memcpy(&len, d + pos, 8);
    memcpy(&len_crc, d + pos + 8, 4);
    if (verify_crc >= 1 && fdio::crc_mask(fdio::crc32c(d + pos, 8)) != len_crc) {
      fdio::fail(FDIO_ERR_CORRUPT, "%s: corrupted record length at byte %zu", path, pos);
      delete f;
      return nullptr;
    }
    if (len > n - pos - 12 || n - pos - 12 - len < 4) {
      fdio::fail(FDIO_ERR_CORRUPT, "%s: truncated record at byte %zu (length %llu)", path, pos, (unsigned long long)len);
      delete f;
      return nullptr;
    }
    if (verify_crc >= 2) {
      uint32_t data_crc;
      memcpy(&data_crc, d + pos + 12 + len, 4);
      if (fdio::crc_mask(fdio::crc32c(d + pos + 12, len)) != data_crc) {
        fdio::fail(FDIO_ERR_CORRUPT, "%s: corrupted record data at byte %zu", path, pos);
        delete f;
        return nullptr;
      }
    }
    f->offset.push_back(pos + 12);
    f->length.push_back(len);
    pos += 12 + len + 4;
  }
  return f;
}

void fdio_tfrecord_close(fdio_tfrecord* f) { delete f; }

int64_t fdio_tfrecord_count(const fdio_tfrecord* f) { return f ? int64_t(f->offset.size()) : 0; }

int fdio_tfrecord_get(const fdio_tfrecord* f, int64_t i, const uint8_t** data, uint64_t* len) {
  if (!f || !data || !len) return fdio::fail(FDIO_ERR_ARG, "fdio_tfrecord_get: null argument");
  if (i < 0 || i >= int64_t(f->offset.size())) return fdio::fail(FDIO_ERR_ARG, "fdio_tfrecord_get: record %lld out of range", (long long)i);
  *data = f->map.data + f->offset[size_t(i)];
  *len = f->length[size_t(i)];
  return FDIO_OK;
}

int fdio_tfrecord_write(const char* path, const uint8_t* const* records, const uint64_t* lens, int64_t n) {
  if (!path || (n > 0 && (!records || !lens))) return fdio::fail(FDIO_ERR_ARG, "fdio_tfrecord_write: null argument");
  FILE* fp = fopen(path, "wb");
  if (!fp) return fdio::fail(FDIO_ERR_IO, "cannot create %s: %s", path, strerror(errno));
  bool ok = true;
  for (int64_t i = 0; i < n && ok; ++i) {
    uint8_t head[12];
    uint64_t len = lens[i];
    memcpy(head, &len, 8);
    uint32_t c = fdio::crc_mask(fdio::crc32c(head, 8));
    memcpy(head + 8, &c, 4);
    uint32_t dc = fdio::crc_mask(fdio::crc32c(records[i], len));
    ok = fwrite(head, 1, 12, fp) == 12 && fwrite(records[i], 1, len, fp) == len && fwrite(&dc, 1, 4, fp) == 4;
  }
  if (fclose(fp) != 0) ok = false;
  return ok ? FDIO_OK : fdio::fail(FDIO_ERR_IO, "short write to %s", path);
}

}  // extern "C"
