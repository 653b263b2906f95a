// TensorFlow tensor-bundle checkpoints: the file format behind Model.save_weights / load_weights of the reference
// (src/mfp/mfp/train.py:67-69,94-97; eval.py:169-172; helpers/callbacks.py:49-56 -> best.ckpt / final.ckpt).
//
// Published formats restated here (TensorFlow itself is not under /root/reference):
//  * <prefix>.index -- tensorflow/core/lib/io/table (the LevelDB table format):
//      data block  = entries { varint32 shared | varint32 non_shared | varint32 value_len | key suffix | value }
//                    + uint32 restart offsets[] + uint32 num_restarts, followed by a 5-byte trailer
//                    { uint8 compression (0 = none, 1 = snappy) | uint32 masked crc32c(block + compression byte) };
//      footer (48 bytes) = BlockHandle metaindex | BlockHandle index | zero padding to 40 | uint64 magic 0xdb4775248b80fb57;
//      BlockHandle = varint64 offset | varint64 size;  index-block values are BlockHandles of the data blocks.
//  * tensorflow/core/protobuf/tensor_bundle.proto:
//      key ""  -> BundleHeaderProto { int32 num_shards = 1; Endianness endianness = 2; VersionDef version = 3; }
//      key k   -> BundleEntryProto  { DataType dtype = 1; TensorShapeProto shape = 2; int32 shard_id = 3; int64 offset = 4;
//                                     int64 size = 5; fixed32 crc32c = 6; repeated TensorSliceProto slices = 7; }
//      TensorShapeProto { repeated Dim dim = 2 { int64 size = 1; string name = 2; }; bool unknown_rank = 3; }
//  * <prefix>.data-SSSSS-of-NNNNN: raw little-endian tensor bytes at [offset, offset + size); entry.crc32c is the masked CRC-32C of them.
//    DT_STRING tensors are stored as varint64 lengths | uint32 masked crc of the lengths | the bytes (read back raw; their CRC recipe
//    differs and is only applied by the writer).
#include <algorithm>
#include <cerrno>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "util.hpp"

namespace {

using fdio::Wire;
constexpr uint64_t kTableMagic = 0xdb4775248b80fb57ull;
constexpr int DT_STRING = 7;

struct Entry {
  std::string key;
  int32_t dtype = 0;
  std::vector<int64_t> dims;
  int32_t shard = 0;
  int64_t offset = 0, size = 0;
  uint32_t crc = 0;
  bool has_slices = false;
};

bool get_handle(Wire* w, uint64_t* off, uint64_t* size) { return w->varint(off) && w->varint(size); }

// Decodes one table block (after checking its trailer) into (key, value) pairs appended to `out`.
int read_block(const fdio::Mapping& m, uint64_t off, uint64_t size, const char* path, std::vector<std::pair<std::string, std::string>>* out) {
  if (off > m.size || size > m.size - off || m.size - off - size < 5) return fdio::fail(FDIO_ERR_CORRUPT, "%s: block handle out of range", path);
  const uint8_t* b = m.data + off;
  const uint8_t type = b[size];
  uint32_t stored;
  memcpy(&stored, b + size + 1, 4);
  if (fdio::crc_mask(fdio::crc32c(b, size + 1)) != stored) return fdio::fail(FDIO_ERR_CORRUPT, "%s: block checksum mismatch at offset %llu", path, (unsigned long long)off);
  if (type != 0) return fdio::fail(FDIO_ERR_UNSUPPORTED, "%s: compressed table block (type %d); tensor bundles are written uncompressed", path, int(type));
  if (size < 4) return fdio::fail(FDIO_ERR_CORRUPT, "%s: block too small", path);
  uint32_t num_restarts;
  memcpy(&num_restarts, b + size - 4, 4);
  if (uint64_t(num_restarts) * 4 + 4 > size) return fdio::fail(FDIO_ERR_CORRUPT, "%s: bad restart array", path);
  const size_t limit = size_t(size) - 4 - size_t(num_restarts) * 4;
  Wire w(b, limit);
  std::string key;
  while (!w.done()) {
    uint64_t shared, non_shared, vlen;
    if (!w.varint(&shared) || !w.varint(&non_shared) || !w.varint(&vlen) || shared > key.size() || non_shared + vlen > size_t(w.end - w.p))
      return fdio::fail(FDIO_ERR_CORRUPT, "%s: bad table entry", path);
    key.resize(size_t(shared));
    key.append(reinterpret_cast<const char*>(w.p), size_t(non_shared));
    w.p += non_shared;
    out->emplace_back(key, std::string(reinterpret_cast<const char*>(w.p), size_t(vlen)));
    w.p += vlen;
  }
  return FDIO_OK;
}

int parse_entry(const std::string& v, Entry* e) {
  Wire w(reinterpret_cast<const uint8_t*>(v.data()), v.size());
  while (!w.done()) {
    uint32_t field, type;
    if (!w.tag(&field, &type)) return FDIO_ERR_CORRUPT;
    uint64_t u;
    if (field == 1 && type == 0) { if (!w.varint(&u)) return FDIO_ERR_CORRUPT; e->dtype = int32_t(u); }
    else if (field == 2 && type == 2) {
      const uint8_t* sp; size_t sn;
      if (!w.bytes(&sp, &sn)) return FDIO_ERR_CORRUPT;
      Wire s(sp, sn);
      while (!s.done()) {
        uint32_t f2, t2;
        if (!s.tag(&f2, &t2)) return FDIO_ERR_CORRUPT;
        if (f2 == 2 && t2 == 2) {
          const uint8_t* dp; size_t dn;
          if (!s.bytes(&dp, &dn)) return FDIO_ERR_CORRUPT;
          Wire d(dp, dn);
          int64_t dim = 0;
          while (!d.done()) {
            uint32_t f3, t3;
            if (!d.tag(&f3, &t3)) return FDIO_ERR_CORRUPT;
            if (f3 == 1 && t3 == 0) { if (!d.varint(&u)) return FDIO_ERR_CORRUPT; dim = int64_t(u); }
            else if (!d.skip(t3)) return FDIO_ERR_CORRUPT;
          }
          e->dims.push_back(dim);
        } else if (!s.skip(t2)) return FDIO_ERR_CORRUPT;
      }
    }
    else if (field == 3 && type == 0) { if (!w.varint(&u)) return FDIO_ERR_CORRUPT; e->shard = int32_t(u); }
    else if (field == 4 && type == 0) { if (!w.varint(&u)) return FDIO_ERR_CORRUPT; e->offset = int64_t(u); }
    else if (field == 5 && type == 0) { if (!w.varint(&u)) return FDIO_ERR_CORRUPT; e->size = int64_t(u); }
    else if (field == 6 && type == 5) { if (!w.fixed32(&e->crc)) return FDIO_ERR_CORRUPT; }
    else { if (field == 7) e->has_slices = true; if (!w.skip(type)) return FDIO_ERR_CORRUPT; }
  }
  return FDIO_OK;
}

std::string shard_name(const std::string& prefix, int shard, int num) {
  char buf[64];
  snprintf(buf, sizeof(buf), ".data-%05d-of-%05d", shard, num);
  return prefix + buf;
}

// ---- table writer -------------------------------------------------------------------------------------------------------------
struct BlockBuilder {
  std::string buf, last;
  std::vector<uint32_t> restarts{0};
  int since_restart = 0;
  void add(const std::string& key, const std::string& value) {
    size_t shared = 0;
    if (since_restart < 16) {
      const size_t m = std::min(last.size(), key.size());
      while (shared < m && last[shared] == key[shared]) ++shared;
    } else {
      restarts.push_back(uint32_t(buf.size()));
      since_restart = 0;
    }
    fdio::put_varint(&buf, shared);
    fdio::put_varint(&buf, key.size() - shared);
    fdio::put_varint(&buf, value.size());
    buf.append(key, shared, std::string::npos);
    buf.append(value);
    last = key;
    ++since_restart;
  }
  bool empty() const { return buf.empty(); }
  std::string finish() {
    std::string out = buf;
    for (uint32_t r : restarts) fdio::put_fixed32(&out, r);
    fdio::put_fixed32(&out, uint32_t(restarts.size()));
    buf.clear(); last.clear(); restarts.assign(1, 0); since_restart = 0;
    return out;
  }
};

std::string handle_bytes(uint64_t off, uint64_t size) {
  std::string s;
  fdio::put_varint(&s, off);
  fdio::put_varint(&s, size);
  return s;
}

// appends block + trailer to `file`; returns its handle encoding
std::string emit_block(std::string* file, const std::string& contents) {
  const uint64_t off = file->size();
  file->append(contents);
  file->push_back('\0');  // no compression
  fdio::put_fixed32(file, fdio::crc_mask(fdio::crc32c(file->data() + off, contents.size() + 1)));
  return handle_bytes(off, contents.size());
}

}  // namespace

struct fdio_bundle {
  std::string prefix;
  int num_shards = 1;
  std::vector<Entry> entries;
  std::vector<std::unique_ptr<fdio::Mapping>> shards;
};

struct fdio_bundle_writer {
  std::string prefix;
  std::map<std::string, Entry> entries;  // sorted by key, as the table requires
  std::string data;
};

extern "C" {

fdio_bundle* fdio_bundle_open(const char* prefix) {
  if (!prefix) { fdio::fail(FDIO_ERR_ARG, "fdio_bundle_open: null prefix"); return nullptr; }
  const std::string index_path = std::string(prefix) + ".index";
  fdio::Mapping m;
  if (m.open(index_path.c_str()) != FDIO_OK) return nullptr;
  const char* path = index_path.c_str();
  if (m.size < 48) { fdio::fail(FDIO_ERR_CORRUPT, "%s: too short to be a table", path); return nullptr; }
  uint64_t magic;
  memcpy(&magic, m.data + m.size - 8, 8);
  if (magic != kTableMagic) { fdio::fail(FDIO_ERR_CORRUPT, "%s: bad table magic number", path); return nullptr; }
  Wire f(m.data + m.size - 48, 40);
  uint64_t meta_off, meta_size, idx_off, idx_size;
  if (!get_handle(&f, &meta_off, &meta_size) || !get_handle(&f, &idx_off, &idx_size)) { fdio::fail(FDIO_ERR_CORRUPT, "%s: bad footer", path); return nullptr; }
  std::vector<std::pair<std::string, std::string>> index, kv;
  if (read_block(m, idx_off, idx_size, path, &index) != FDIO_OK) return nullptr;
  for (auto& iv : index) {
    Wire h(reinterpret_cast<const uint8_t*>(iv.second.data()), iv.second.size());
    uint64_t off, size;
    if (!get_handle(&h, &off, &size)) { fdio::fail(FDIO_ERR_CORRUPT, "%s: bad index entry", path); return nullptr; }
    if (read_block(m, off, size, path, &kv) != FDIO_OK) return nullptr;
  }
  auto b = std::make_unique<fdio_bundle>();
  b->prefix = prefix;
  bool header = false;
  for (auto& e : kv) {
    if (e.first.empty()) {  // BundleHeaderProto
      Wire w(reinterpret_cast<const uint8_t*>(e.second.data()), e.second.size());
      while (!w.done()) {
        uint32_t field, type; uint64_t u;
        if (!w.tag(&field, &type)) { fdio::fail(FDIO_ERR_CORRUPT, "%s: bad bundle header", path); return nullptr; }
        if (field == 1 && type == 0) { if (!w.varint(&u)) { fdio::fail(FDIO_ERR_CORRUPT, "%s: bad bundle header", path); return nullptr; } b->num_shards = int(u); }
        else if (field == 2 && type == 0) {
          if (!w.varint(&u)) { fdio::fail(FDIO_ERR_CORRUPT, "%s: bad bundle header", path); return nullptr; }
          if (u != 0) { fdio::fail(FDIO_ERR_UNSUPPORTED, "%s: big-endian bundle", path); return nullptr; }
        } else if (!w.skip(type)) { fdio::fail(FDIO_ERR_CORRUPT, "%s: bad bundle header", path); return nullptr; }
      }
      header = true;
      continue;
    }
    Entry en;
    en.key = e.first;
    if (parse_entry(e.second, &en) != FDIO_OK) { fdio::fail(FDIO_ERR_CORRUPT, "%s: bad entry for key '%s'", path, e.first.c_str()); return nullptr; }
    b->entries.push_back(std::move(en));
  }
  if (!header) { fdio::fail(FDIO_ERR_CORRUPT, "%s: no bundle header (key \"\")", path); return nullptr; }
  if (b->num_shards < 1 || b->num_shards > 99999) { fdio::fail(FDIO_ERR_CORRUPT, "%s: bad shard count %d", path, b->num_shards); return nullptr; }
  if (!std::is_sorted(b->entries.begin(), b->entries.end(), [](const Entry& x, const Entry& y) { return x.key < y.key; })) {
    fdio::fail(FDIO_ERR_CORRUPT, "%s: keys are not sorted", path);
    return nullptr;
  }
  b->shards.resize(size_t(b->num_shards));
  return b.release();
}

void fdio_bundle_close(fdio_bundle* b) { delete b; }
int32_t fdio_bundle_count(const fdio_bundle* b) { return b ? int32_t(b->entries.size()) : 0; }
const char* fdio_bundle_key(const fdio_bundle* b, int32_t i) {
  if (!b || i < 0 || i >= int32_t(b->entries.size())) { fdio::fail(FDIO_ERR_ARG, "fdio_bundle_key: index out of range"); return nullptr; }
  return b->entries[size_t(i)].key.c_str();
}
int32_t fdio_bundle_find(const fdio_bundle* b, const char* key) {
  if (!b || !key) return fdio::fail(FDIO_ERR_ARG, "fdio_bundle_find: null argument");
  auto it = std::lower_bound(b->entries.begin(), b->entries.end(), std::string(key), [](const Entry& e, const std::string& k) { return e.key < k; });
  if (it == b->entries.end() || it->key != key) return fdio::fail(FDIO_ERR_NOT_FOUND, "key '%s' is not in the checkpoint", key);
  return int32_t(it - b->entries.begin());
}
int fdio_bundle_info(const fdio_bundle* b, int32_t i, int32_t* dtype, int32_t* rank, int64_t* dims, int32_t max_rank, int64_t* nbytes) {
  if (!b || i < 0 || i >= int32_t(b->entries.size())) return fdio::fail(FDIO_ERR_ARG, "fdio_bundle_info: index out of range");
  const Entry& e = b->entries[size_t(i)];
  if (dtype) *dtype = e.dtype;
  if (rank) *rank = int32_t(e.dims.size());
  if (dims) for (int32_t k = 0; k < max_rank && k < int32_t(e.dims.size()); ++k) dims[k] = e.dims[size_t(k)];
  if (nbytes) *nbytes = e.size;
  return FDIO_OK;
}
int fdio_bundle_read(const fdio_bundle* cb, int32_t i, void* dst, int64_t nbytes) {
  auto* b = const_cast<fdio_bundle*>(cb);  // shard mappings are opened on first use
  if (!b || i < 0 || i >= int32_t(b->entries.size()) || (!dst && nbytes > 0)) return fdio::fail(FDIO_ERR_ARG, "fdio_bundle_read: bad argument");
  const Entry& e = b->entries[size_t(i)];
  if (e.has_slices) return fdio::fail(FDIO_ERR_UNSUPPORTED, "key '%s' is a partitioned (sliced) variable", e.key.c_str());
  if (nbytes != e.size) return fdio::fail(FDIO_ERR_ARG, "key '%s': buffer of %lld bytes for a tensor of %lld", e.key.c_str(), (long long)nbytes, (long long)e.size);
  if (e.shard < 0 || e.shard >= b->num_shards) return fdio::fail(FDIO_ERR_CORRUPT, "key '%s': shard %d of %d", e.key.c_str(), e.shard, b->num_shards);
  auto& m = b->shards[size_t(e.shard)];
  if (!m) {
    auto mm = std::make_unique<fdio::Mapping>();
    if (mm->open(shard_name(b->prefix, e.shard, b->num_shards).c_str()) != FDIO_OK) return FDIO_ERR_IO;
    m = std::move(mm);
  }
  if (e.offset < 0 || e.size < 0 || uint64_t(e.offset) > m->size || uint64_t(e.size) > m->size - uint64_t(e.offset))
    return fdio::fail(FDIO_ERR_CORRUPT, "key '%s': bytes [%lld, +%lld) are outside the data shard", e.key.c_str(), (long long)e.offset, (long long)e.size);
  const uint8_t* src = m->data + e.offset;
  if (e.dtype != DT_STRING && fdio::crc_mask(fdio::crc32c(src, size_t(e.size))) != e.crc)
    return fdio::fail(FDIO_ERR_CORRUPT, "key '%s': tensor checksum mismatch", e.key.c_str());
  if (e.size) memcpy(dst, src, size_t(e.size));
  return FDIO_OK;
}

fdio_bundle_writer* fdio_bundle_writer_create(const char* prefix) {
  if (!prefix) { fdio::fail(FDIO_ERR_ARG, "fdio_bundle_writer_create: null prefix"); return nullptr; }
  auto* w = new fdio_bundle_writer;
  w->prefix = prefix;
  return w;
}

int fdio_bundle_writer_add(fdio_bundle_writer* w, const char* key, int32_t dtype, int32_t rank, const int64_t* dims, const void* data, int64_t nbytes) {
  if (!w || !key || !*key || rank < 0 || (rank && !dims) || nbytes < 0 || (nbytes && !data)) return fdio::fail(FDIO_ERR_ARG, "fdio_bundle_writer_add: bad argument");
  if (w->entries.count(key)) return fdio::fail(FDIO_ERR_ARG, "key '%s' added twice", key);
  Entry e;
  e.key = key;
  e.dtype = dtype;
  e.dims.assign(dims, dims + rank);
  e.offset = int64_t(w->data.size());
  const uint8_t* src = static_cast<const uint8_t*>(data);
  if (dtype == DT_STRING) {
    // caller passes the concatenation of the elements preceded by nothing; only scalar strings are written (the object graph)
    if (rank != 0) return fdio::fail(FDIO_ERR_UNSUPPORTED, "only scalar DT_STRING tensors can be written");
    std::string lengths;
    fdio::put_varint(&lengths, uint64_t(nbytes));
    uint32_t crc = 0;
    const uint32_t len32 = uint32_t(nbytes);
    crc = fdio::crc32c_extend(crc, reinterpret_cast<const uint8_t*>(&len32), 4);
    const uint32_t length_checksum = fdio::crc_mask(crc);
    w->data.append(lengths);
    fdio::put_fixed32(&w->data, length_checksum);
    crc = fdio::crc32c_extend(crc, reinterpret_cast<const uint8_t*>(&length_checksum), 4);
    w->data.append(reinterpret_cast<const char*>(src), size_t(nbytes));
    crc = fdio::crc32c_extend(crc, src, size_t(nbytes));
    e.crc = fdio::crc_mask(crc);
    e.size = int64_t(w->data.size()) - e.offset;
  } else {
    w->data.append(reinterpret_cast<const char*>(src), size_t(nbytes));
    e.crc = fdio::crc_mask(fdio::crc32c(src, size_t(nbytes)));
    e.size = nbytes;
  }
  w->entries.emplace(e.key, std::move(e));
  return FDIO_OK;
}

int fdio_bundle_writer_finish(fdio_bundle_writer* wr) {
  if (!wr) return fdio::fail(FDIO_ERR_ARG, "fdio_bundle_writer_finish: null writer");
  std::unique_ptr<fdio_bundle_writer> w(wr);
  auto write_file = [](const std::string& path, const std::string& bytes) {
    FILE* fp = fopen(path.c_str(), "wb");
    if (!fp) return fdio::fail(FDIO_ERR_IO, "cannot create %s: %s", path.c_str(), strerror(errno));
    bool ok = fwrite(bytes.data(), 1, bytes.size(), fp) == bytes.size();
    if (fclose(fp) != 0) ok = false;
    return ok ? int(FDIO_OK) : fdio::fail(FDIO_ERR_IO, "short write to %s", path.c_str());
  };
  int code = write_file(shard_name(w->prefix, 0, 1), w->data);
  if (code != FDIO_OK) return code;

  std::string file;
  BlockBuilder data_block, index_block;
  std::string last_key;
  auto flush = [&] {
    if (data_block.empty()) return;
    std::string h = emit_block(&file, data_block.finish());
    index_block.add(last_key, h);  // any key >= the block's last key and < the next block's first separates them
  };
  auto add = [&](const std::string& key, const std::string& value) {
    data_block.add(key, value);
    last_key = key;
    if (data_block.buf.size() >= 4096) flush();
  };
  std::string header;  // BundleHeaderProto { num_shards = 1, endianness = LITTLE (0, default: omitted), version { producer = 1 } }
  fdio::put_tag(&header, 1, 0); fdio::put_varint(&header, 1);
  { std::string ver; fdio::put_tag(&ver, 1, 0); fdio::put_varint(&ver, 1); fdio::put_bytes(&header, 3, ver); }
  add("", header);
  for (auto& kv : w->entries) {
    const Entry& e = kv.second;
    std::string v, shape;
    if (e.dtype) { fdio::put_tag(&v, 1, 0); fdio::put_varint(&v, uint64_t(e.dtype)); }
    for (int64_t d : e.dims) { std::string dim; if (d) { fdio::put_tag(&dim, 1, 0); fdio::put_varint(&dim, uint64_t(d)); } fdio::put_bytes(&shape, 2, dim); }
    fdio::put_bytes(&v, 2, shape);
    if (e.offset) { fdio::put_tag(&v, 4, 0); fdio::put_varint(&v, uint64_t(e.offset)); }
    if (e.size) { fdio::put_tag(&v, 5, 0); fdio::put_varint(&v, uint64_t(e.size)); }
    fdio::put_tag(&v, 6, 5); fdio::put_fixed32(&v, e.crc);
    add(e.key, v);
  }
  flush();
  BlockBuilder meta;
  std::string meta_handle = emit_block(&file, meta.finish());
  std::string index_handle = emit_block(&file, index_block.finish());
  std::string footer = meta_handle + index_handle;
  footer.resize(40, '\0');
  fdio::put_fixed64(&footer, kTableMagic);
  file.append(footer);
  return write_file(w->prefix + ".index", file);
}

}  // extern "C"
