// tf.train.SequenceExample -> padded batch columns: the native side of DataSpec.parse_fn
// (src/mfp/mfp/data/spec.py:255-287) with the preprocessors of DataSpec._init_preprocessor (spec.py:90-134) and
// SequenceDiscretizer (data/discretizer.py:6-31) applied while the values are written.
//
// Published schema (tensorflow/core/example/{example,feature}.proto):
//   SequenceExample { Features context = 1; FeatureLists feature_lists = 2; }
//   Features        { map<string, Feature> feature = 1; }          FeatureLists { map<string, FeatureList> feature_list = 1; }
//   FeatureList     { repeated Feature feature = 1; }
//   Feature         { oneof kind { BytesList bytes_list = 1; FloatList float_list = 2; Int64List int64_list = 3; } }
//   BytesList { repeated bytes value = 1; }  FloatList { repeated float value = 1 [packed]; }  Int64List { repeated int64 value = 1 [packed]; }
// Semantics restated from tf.io.parse_sequence_example with FixedLenFeature / FixedLenSequenceFeature (no defaults, allow_missing
// False): a missing key, a wrong kind or a value count != prod(shape) is an error; sequence features are padded to the longest of the
// batch with 0 / 0.0 / "".
#include <algorithm>
#include <atomic>
#include <memory>
#include <string>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <pthread.h>
#include <thread>
#include <unordered_map>
#include <vector>

#include "util.hpp"

namespace {

using fdio::Wire;

// Vocabulary of a string lookup: open addressing over FNV-1a, probed with the record's bytes in place (no temporary std::string).
struct StringTable {
  std::vector<int32_t> slot;  // index into keys, -1 = empty; size = power of two >= 2 * keys
  std::vector<std::string> keys;
  std::vector<int32_t> values;
  static uint64_t hash(const uint8_t* p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h ^ (h >> 29);
  }
  bool insert(const std::string& key, int32_t value) {
    if (find(reinterpret_cast<const uint8_t*>(key.data()), key.size()) != nullptr) return false;
    keys.push_back(key);
    values.push_back(value);
    if (slot.size() < 2 * keys.size() + 2) {
      size_t cap = 16;
      while (cap < 4 * keys.size()) cap <<= 1;
      slot.assign(cap, -1);
      for (size_t k = 0; k < keys.size(); ++k) place(int32_t(k));
    } else {
      place(int32_t(keys.size() - 1));
    }
    return true;
  }
  void place(int32_t k) {
    size_t i = hash(reinterpret_cast<const uint8_t*>(keys[size_t(k)].data()), keys[size_t(k)].size()) & (slot.size() - 1);
    while (slot[i] >= 0) i = (i + 1) & (slot.size() - 1);
    slot[i] = k;
  }
  const int32_t* find(const uint8_t* p, size_t n) const {
    if (slot.empty()) return nullptr;
    size_t i = hash(p, n) & (slot.size() - 1);
    for (int32_t k; (k = slot[i]) >= 0; i = (i + 1) & (slot.size() - 1)) {
      const std::string& key = keys[size_t(k)];
      if (key.size() == n && memcmp(key.data(), p, n) == 0) return &values[size_t(k)];
    }
    return nullptr;
  }
};

struct Column {
  std::string name;
  int is_sequence, dtype, width, transform, output;
  int num_oov, has_mask;
  std::string mask_str;
  int64_t mask_int;
  StringTable str_index;
  std::unordered_map<int64_t, int32_t> int_index;
  std::vector<float> boundaries;
  int32_t first_vocab_index;  // [mask] + [oov] come first
  // value a padded step holds after preprocessing; pad_code != OK when the parse default is not representable (OOV without OOV index)
  int32_t pad_i32 = 0;
  int pad_code = FDIO_OK;
};

struct Span { const uint8_t* p = nullptr; size_t n = 0; bool present = false; };

static const char* kDtypeName[] = {"int64", "float", "string"};

struct Value {  // one scalar as it sits in the record
  int64_t i = 0;
  float f = 0.f;
  const uint8_t* s = nullptr;
  size_t n = 0;
};

}  // namespace

struct fdio_schema {
  std::vector<Column> cols;
  std::unordered_map<std::string, int> context_index, sequence_index;
};

namespace {

// ---- preprocessors ----------------------------------------------------------------------------------------------------------------
int lookup_str(const Column& c, const uint8_t* s, size_t n, int32_t* out) {
  if (c.has_mask && n == c.mask_str.size() && memcmp(s, c.mask_str.data(), n) == 0) { *out = 0; return FDIO_OK; }
  if (const int32_t* hit = c.str_index.find(s, n)) { *out = *hit; return FDIO_OK; }
  if (c.num_oov == 1) { *out = c.has_mask ? 1 : 0; return FDIO_OK; }
  return FDIO_ERR_OOV;
}
int lookup_int(const Column& c, int64_t v, int32_t* out) {
  if (c.has_mask && v == c.mask_int) { *out = 0; return FDIO_OK; }
  auto it = c.int_index.find(v);
  if (it != c.int_index.end()) { *out = it->second; return FDIO_OK; }
  if (c.num_oov == 1) { *out = c.has_mask ? 1 : 0; return FDIO_OK; }
  return FDIO_ERR_OOV;
}
inline int32_t bucketize(const Column& c, float x) {  // Bucketize: number of boundaries <= x
  return int32_t(std::upper_bound(c.boundaries.begin(), c.boundaries.end(), x) - c.boundaries.begin());
}

// One value -> its place in the output buffer.  `slot` counts scalars from the start of the column buffer.
int emit(const Column& c, const Value& v, void* out, size_t slot, const uint8_t* record) {
  if (c.output == FDIO_OUT_SKIP) {
    if (c.transform != FDIO_LOOKUP) return FDIO_OK;  // still validate lookups so that errors do not depend on the output kind
  }
  int32_t r = 0;
  switch (c.transform) {
    case FDIO_LOOKUP: {
      int code = c.dtype == FDIO_STRING ? lookup_str(c, v.s, v.n, &r) : lookup_int(c, v.i, &r);
      if (code != FDIO_OK) return code;
      break;
    }
    case FDIO_DISCRETIZE:
      r = bucketize(c, c.dtype == FDIO_INT64 ? float(v.i) : v.f);
      break;
    default:
      if (c.dtype == FDIO_FLOAT32) {
        if (c.output == FDIO_OUT_FLOAT32) static_cast<float*>(out)[slot] = v.f;
        return FDIO_OK;
      }
      if (c.dtype == FDIO_STRING) {
        if (c.output == FDIO_OUT_SPAN) {
          int64_t* o = static_cast<int64_t*>(out) + 2 * slot;
          o[0] = v.s ? int64_t(v.s - record) : 0;
          o[1] = int64_t(v.n);
        }
        return FDIO_OK;
      }
      r = int32_t(v.i);  // tf.cast(int64 -> int32), spec.py:281-285
  }
  if (c.output == FDIO_OUT_INT32) static_cast<int32_t*>(out)[slot] = r;
  return FDIO_OK;
}

// ---- Feature decoding -----------------------------------------------------------------------------------------------------------
// Walks the values of one Feature message; calls fn(Value) per scalar.  Returns the count or a negative code.
// `bulk` (optional): destination of `bulk_n` untransformed floats -- a packed FloatList of exactly that many values is copied in one go.
template <class Fn>
int64_t for_each_value(const Column& c, const uint8_t* p, size_t n, Fn&& fn, float* bulk = nullptr, size_t bulk_n = 0) {
  Wire w(p, n);
  int64_t count = 0;
  bool kind_seen = false;
  while (!w.done()) {
    uint32_t field, type;
    if (!w.tag(&field, &type)) return FDIO_ERR_CORRUPT;
    if (type != 2 || field < 1 || field > 3) { if (!w.skip(type)) return FDIO_ERR_CORRUPT; continue; }
    const uint8_t* lp; size_t ln;
    if (!w.bytes(&lp, &ln)) return FDIO_ERR_CORRUPT;
    const int kind = field == 1 ? FDIO_STRING : field == 2 ? FDIO_FLOAT32 : FDIO_INT64;
    if (kind != c.dtype) return FDIO_ERR_INVALID;
    if (kind_seen) count = 0;  // oneof: the last kind on the wire wins; the callers below only see the final pass
    kind_seen = true;
    Wire l(lp, ln);
    while (!l.done()) {
      uint32_t f2, t2;
      if (!l.tag(&f2, &t2)) return FDIO_ERR_CORRUPT;
      if (f2 != 1) { if (!l.skip(t2)) return FDIO_ERR_CORRUPT; continue; }
      Value v;
      if (kind == FDIO_STRING) {
        if (t2 != 2 || !l.bytes(&v.s, &v.n)) return FDIO_ERR_CORRUPT;
        int code = fn(v, count); if (code != FDIO_OK) return code;
        ++count;
      } else if (kind == FDIO_FLOAT32) {
        if (t2 == 2) {  // packed
          const uint8_t* fp; size_t fn_;
          if (!l.bytes(&fp, &fn_) || (fn_ & 3)) return FDIO_ERR_CORRUPT;
          if (bulk && count == 0 && fn_ == 4 * bulk_n) { memcpy(bulk, fp, fn_); count += int64_t(bulk_n); continue; }
          for (size_t k = 0; k < fn_; k += 4) { memcpy(&v.f, fp + k, 4); int code = fn(v, count); if (code != FDIO_OK) return code; ++count; }
        } else if (t2 == 5) {
          uint32_t bits; if (!l.fixed32(&bits)) return FDIO_ERR_CORRUPT;
          memcpy(&v.f, &bits, 4);
          int code = fn(v, count); if (code != FDIO_OK) return code;
          ++count;
        } else return FDIO_ERR_CORRUPT;
      } else {
        if (t2 == 2) {  // packed varints
          const uint8_t* ip; size_t in_;
          if (!l.bytes(&ip, &in_)) return FDIO_ERR_CORRUPT;
          Wire iv(ip, in_);
          while (!iv.done()) { uint64_t u; if (!iv.varint(&u)) return FDIO_ERR_CORRUPT; v.i = int64_t(u); int code = fn(v, count); if (code != FDIO_OK) return code; ++count; }
        } else if (t2 == 0) {
          uint64_t u; if (!l.varint(&u)) return FDIO_ERR_CORRUPT;
          v.i = int64_t(u);
          int code = fn(v, count); if (code != FDIO_OK) return code;
          ++count;
        } else return FDIO_ERR_CORRUPT;
      }
    }
  }
  return kind_seen ? count : int64_t(FDIO_ERR_NOT_FOUND);  // a Feature with no kind set holds no values
}

// The canonical encodings of a Feature holding exactly one value (what every scalar column of a step looks like; all lengths < 128):
//   bytes : 0a L+2 0a L <bytes>        float : 12 06 0a 04 <f32> (packed) or 12 05 0d <f32>        int64 : 1a L+2 0a L <varint> or 1a L+1 08 <varint>
// Returns true and fills v when the bytes have exactly that shape and the kind is the column's; anything else goes through
// for_each_value, which handles every legal encoding and reports the errors.
inline bool single_value(const Column& c, const uint8_t* p, size_t n, Value* v) {
  if (n < 3 || n > 129 || p[1] != n - 2) return false;
  if (p[0] == 0x0a && c.dtype == FDIO_STRING) {
    if (n < 4 || p[2] != 0x0a || p[3] != n - 4) return false;
    v->s = p + 4; v->n = n - 4;
    return true;
  }
  if (p[0] == 0x12 && c.dtype == FDIO_FLOAT32) {
    if (n == 8 && p[2] == 0x0a && p[3] == 0x04) { memcpy(&v->f, p + 4, 4); return true; }
    if (n == 7 && p[2] == 0x0d) { memcpy(&v->f, p + 3, 4); return true; }
    return false;
  }
  if (p[0] == 0x1a && c.dtype == FDIO_INT64) {
    size_t at;
    if (p[2] == 0x0a) { if (n < 5 || p[3] != n - 4) return false; at = 4; }
    else if (p[2] == 0x08) at = 3;
    else return false;
    uint64_t u = 0;
    int shift = 0;
    for (; at < n && shift < 64; shift += 7) {
      const uint8_t byte = p[at++];
      u |= uint64_t(byte & 0x7f) << shift;
      if (!(byte & 0x80)) { if (at != n) return false; v->i = int64_t(u); return true; }
    }
    return false;
  }
  return false;
}

// Collects the (key -> message bytes) entries of a Features / FeatureLists map; last occurrence of a key wins (protobuf map rule).
int collect_map(const uint8_t* p, size_t n, const std::unordered_map<std::string, int>& index, std::vector<Span>* found) {
  Wire w(p, n);
  std::string key;
  while (!w.done()) {
    uint32_t field, type;
    if (!w.tag(&field, &type)) return FDIO_ERR_CORRUPT;
    if (field != 1 || type != 2) { if (!w.skip(type)) return FDIO_ERR_CORRUPT; continue; }
    const uint8_t* ep; size_t en;
    if (!w.bytes(&ep, &en)) return FDIO_ERR_CORRUPT;
    Wire e(ep, en);
    const uint8_t* kp = nullptr; size_t kn = 0;
    Span val; val.present = true;  // an entry without a value field is an empty message
    while (!e.done()) {
      uint32_t f2, t2;
      if (!e.tag(&f2, &t2)) return FDIO_ERR_CORRUPT;
      if (f2 == 1 && t2 == 2) { if (!e.bytes(&kp, &kn)) return FDIO_ERR_CORRUPT; }
      else if (f2 == 2 && t2 == 2) { if (!e.bytes(&val.p, &val.n)) return FDIO_ERR_CORRUPT; }
      else if (!e.skip(t2)) return FDIO_ERR_CORRUPT;
    }
    key.assign(reinterpret_cast<const char*>(kp), kn);
    auto it = index.find(key);
    if (it != index.end()) (*found)[size_t(it->second)] = val;
  }
  return FDIO_OK;
}

int split_example(const uint8_t* rec, size_t len, const fdio_schema& s, std::vector<Span>* found) {
  std::fill(found->begin(), found->end(), Span());
  Wire w(rec, len);
  while (!w.done()) {
    uint32_t field, type;
    if (!w.tag(&field, &type)) return FDIO_ERR_CORRUPT;
    if (type == 2 && (field == 1 || field == 2)) {
      const uint8_t* p; size_t n;
      if (!w.bytes(&p, &n)) return FDIO_ERR_CORRUPT;
      int code = collect_map(p, n, field == 1 ? s.context_index : s.sequence_index, found);
      if (code != FDIO_OK) return code;
    } else if (!w.skip(type)) return FDIO_ERR_CORRUPT;
  }
  return FDIO_OK;
}

// Number of Feature entries of a FeatureList.
int64_t count_steps(const Span& fl) {
  Wire w(fl.p, fl.n);
  int64_t steps = 0;
  while (!w.done()) {
    uint32_t field, type;
    if (!w.tag(&field, &type)) return FDIO_ERR_CORRUPT;
    if (field == 1 && type == 2) { const uint8_t* p; size_t n; if (!w.bytes(&p, &n)) return FDIO_ERR_CORRUPT; ++steps; }
    else if (!w.skip(type)) return FDIO_ERR_CORRUPT;
  }
  return steps;
}

struct RecordError { int code = FDIO_OK; std::string text; };

void describe(RecordError* e, int code, int32_t b, const Column& c, int64_t step, const char* what) {
  char buf[400];
  if (step >= 0) snprintf(buf, sizeof(buf), "record %d, key '%s', index %lld: %s", b, c.name.c_str(), (long long)step, what);
  else snprintf(buf, sizeof(buf), "record %d, key '%s': %s", b, c.name.c_str(), what);
  e->code = code;
  e->text = buf;
}

const char* what_for(int code, const Column& c, char* scratch, size_t n) {
  switch (code) {
    case FDIO_ERR_CORRUPT: return "malformed protobuf";
    case FDIO_ERR_INVALID: snprintf(scratch, n, "feature kind does not match the column dtype (%s)", kDtypeName[c.dtype]); return scratch;
    case FDIO_ERR_OOV: return "value is not in the lookup vocabulary and num_oov_indices is 0";
    case FDIO_ERR_NOT_FOUND: return "feature holds no values";
    default: return "error";
  }
}

// skip (optional, one flag per column): columns a later pass writes (the packed columns of fdio_parse_batch_packed)
void parse_record(const fdio_schema& s, const uint8_t* rec, size_t len, int32_t b, int32_t S, void* const* out, std::vector<Span>* found,
                  RecordError* err, const uint8_t* skip = nullptr) {
  char scratch[160];
  int code = split_example(rec, len, s, found);
  if (code != FDIO_OK) { err->code = code; err->text = "record " + std::to_string(b) + ": malformed SequenceExample"; return; }
  for (size_t ci = 0; ci < s.cols.size(); ++ci) {
    if (skip && skip[ci]) continue;
    const Column& c = s.cols[ci];
    const Span& sp = (*found)[ci];
    void* o = out ? out[ci] : nullptr;
    const size_t W = size_t(c.width);
    if (!c.is_sequence) {
      if (!sp.present) { describe(err, FDIO_ERR_INVALID, b, c, -1, "feature is required but could not be found"); return; }
      const size_t base = size_t(b) * W;
      int64_t n = for_each_value(c, sp.p, sp.n, [&](const Value& v, int64_t k) { return k < int64_t(W) ? emit(c, v, o, base + size_t(k), rec) : FDIO_OK; });
      if (n < 0) { describe(err, n == FDIO_ERR_NOT_FOUND ? int(FDIO_ERR_INVALID) : int(n), b, c, -1, what_for(int(n), c, scratch, sizeof(scratch))); return; }
      if (n != int64_t(W)) {
        snprintf(scratch, sizeof(scratch), "number of %s values != expected: values size %lld but output shape holds %zu", kDtypeName[c.dtype], (long long)n, W);
        describe(err, FDIO_ERR_INVALID, b, c, -1, scratch);
        return;
      }
      continue;
    }
    if (!sp.present) { describe(err, FDIO_ERR_INVALID, b, c, -1, "feature list is required but could not be found"); return; }
    // steps of this document
    Wire w(sp.p, sp.n);
    int64_t t = 0;
    while (!w.done()) {
      uint32_t field, type;
      if (!w.tag(&field, &type)) { describe(err, FDIO_ERR_CORRUPT, b, c, t, "malformed protobuf"); return; }
      if (field != 1 || type != 2) { if (!w.skip(type)) { describe(err, FDIO_ERR_CORRUPT, b, c, t, "malformed protobuf"); return; } continue; }
      const uint8_t* fp; size_t fn;
      if (!w.bytes(&fp, &fn)) { describe(err, FDIO_ERR_CORRUPT, b, c, t, "malformed protobuf"); return; }
      if (t >= S) { describe(err, FDIO_ERR_ARG, b, c, t, "more steps than the batch was sized for"); return; }
      const size_t base = (size_t(b) * size_t(S) + size_t(t)) * W;
      Value one;
      if (W == 1 && single_value(c, fp, fn, &one)) {
        const int code = emit(c, one, o, base, rec);
        if (code != FDIO_OK) { describe(err, code, b, c, t, what_for(code, c, scratch, sizeof(scratch))); return; }
        ++t;
        continue;
      }
      float* bulk = (c.transform == FDIO_NONE && c.output == FDIO_OUT_FLOAT32) ? static_cast<float*>(o) + base : nullptr;
      int64_t n = for_each_value(c, fp, fn, [&](const Value& v, int64_t k) { return k < int64_t(W) ? emit(c, v, o, base + size_t(k), rec) : FDIO_OK; }, bulk, W);
      if (n < 0) { describe(err, n == FDIO_ERR_NOT_FOUND ? int(FDIO_ERR_INVALID) : int(n), b, c, t, what_for(int(n), c, scratch, sizeof(scratch))); return; }
      if (n != int64_t(W)) {
        snprintf(scratch, sizeof(scratch), "number of %s values != expected: values size %lld but output shape holds %zu", kDtypeName[c.dtype], (long long)n, W);
        describe(err, FDIO_ERR_INVALID, b, c, t, scratch);
        return;
      }
      ++t;
    }
    // padding: the parse default put through the preprocessor
    if (t < S && c.output != FDIO_OUT_SKIP) {
      if (c.pad_code != FDIO_OK) { describe(err, c.pad_code, b, c, t, "the padding value is not in the lookup vocabulary and num_oov_indices is 0"); return; }
      const size_t base = (size_t(b) * size_t(S) + size_t(t)) * W, cnt = size_t(S - t) * W;
      if (c.output == FDIO_OUT_INT32) std::fill_n(static_cast<int32_t*>(o) + base, cnt, c.pad_i32);
      else if (c.output == FDIO_OUT_FLOAT32) std::fill_n(static_cast<float*>(o) + base, cnt, 0.f);
      else std::fill_n(static_cast<int64_t*>(o) + 2 * base, 2 * cnt, int64_t(0));
    }
  }
}

// Does element (b, t) carry packed column pc?  (fdio_parse_batch_packed; the length and gate columns are already in `out`)
inline int carries(const fdio_schema& s, const fdio_pack_column& pc, void* const* out, int32_t length_column, int32_t b, int32_t S, int32_t t) {
  if (t > static_cast<const int32_t*>(out[length_column])[size_t(b) * size_t(s.cols[size_t(length_column)].width)]) return 0;
  if (pc.cond_column < 0) return 1;
  const int32_t v = static_cast<const int32_t*>(out[pc.cond_column])[size_t(b) * size_t(S) + size_t(t)];
  if (v < 0 || v >= pc.cond_n) return -1;
  return pc.cond_mask[v] ? 1 : 0;
}

// Is this Feature exactly one packed FloatList of W values in the canonical encoding -- 12 <len> 0a <4W> <4W bytes> -- ?  Decided from
// the few header bytes and the lengths alone: the payload of a row nothing will read is then not touched at all (the records are
// memory-mapped: 59 % of crello's embedding bytes never leave the page cache).  Anything else is decoded in full by for_each_value.
inline bool is_plain_float_row(const uint8_t* p, size_t n, size_t W) {
  Wire w(p, n);
  uint32_t field, type;
  const uint8_t* lp; size_t ln;
  if (!w.tag(&field, &type) || field != 2 || type != 2 || !w.bytes(&lp, &ln) || !w.done()) return false;
  Wire l(lp, ln);
  const uint8_t* fp; size_t fn;
  if (!l.tag(&field, &type) || field != 1 || type != 2 || !l.bytes(&fp, &fn) || !l.done()) return false;
  return fn == 4 * W;
}

// One packed column of one record: rows of the elements that carry it go to out[column] from row `row0` on, the row map gets the row or -1.
void pack_record(const fdio_schema& s, const fdio_pack_column& pc, const Span& sp, int32_t b, int32_t S, void* const* out,
                 int32_t length_column, int64_t row0, float* scratch_row, RecordError* err) {
  char scratch[160];
  const Column& c = s.cols[size_t(pc.column)];
  const size_t W = size_t(c.width);
  float* dst = static_cast<float*>(out[pc.column]);
  int32_t* map = pc.rowmap + size_t(b) * size_t(S);
  if (!sp.present) { describe(err, FDIO_ERR_INVALID, b, c, -1, "feature list is required but could not be found"); return; }
  Wire w(sp.p, sp.n);
  int64_t t = 0, row = row0;
  while (!w.done()) {
    uint32_t field, type;
    if (!w.tag(&field, &type)) { describe(err, FDIO_ERR_CORRUPT, b, c, t, "malformed protobuf"); return; }
    if (field != 1 || type != 2) { if (!w.skip(type)) { describe(err, FDIO_ERR_CORRUPT, b, c, t, "malformed protobuf"); return; } continue; }
    const uint8_t* fp; size_t fn;
    if (!w.bytes(&fp, &fn)) { describe(err, FDIO_ERR_CORRUPT, b, c, t, "malformed protobuf"); return; }
    if (t >= S) { describe(err, FDIO_ERR_ARG, b, c, t, "more steps than the batch was sized for"); return; }
    const int carry = carries(s, pc, out, length_column, b, S, int32_t(t));
    if (carry < 0) { describe(err, FDIO_ERR_ARG, b, c, t, "the gating column's value is outside the loss_condition mask"); return; }
    if (!carry && is_plain_float_row(fp, fn, W)) { map[t] = -1; ++t; continue; }  // well-formed and unread: not even decoded
    float* row_dst = carry ? dst + size_t(row) * W : scratch_row;  // other encodings of a row nothing reads are decoded all the same (same errors as the dense parse)
    int64_t n = for_each_value(c, fp, fn, [&](const Value& v, int64_t k) { if (k < int64_t(W)) row_dst[k] = v.f; return int(FDIO_OK); }, row_dst, W);
    if (n < 0) { describe(err, n == FDIO_ERR_NOT_FOUND ? int(FDIO_ERR_INVALID) : int(n), b, c, t, what_for(int(n), c, scratch, sizeof(scratch))); return; }
    if (n != int64_t(W)) {
      snprintf(scratch, sizeof(scratch), "number of %s values != expected: values size %lld but output shape holds %zu", kDtypeName[c.dtype], (long long)n, W);
      describe(err, FDIO_ERR_INVALID, b, c, t, scratch);
      return;
    }
    map[t] = carry ? int32_t(row++) : -1;
    ++t;
  }
  for (; t < S; ++t) {  // past the record's own steps: the dense parse holds zeros there, and so does a row an inconsistent length asks for
    const int carry = carries(s, pc, out, length_column, b, S, int32_t(t));
    if (carry < 0) { describe(err, FDIO_ERR_ARG, b, c, t, "the gating column's value is outside the loss_condition mask"); return; }
    if (carry) { std::fill_n(dst + size_t(row) * W, W, 0.f); map[t] = int32_t(row++); }
    else map[t] = -1;
  }
}

// Worker pool of the parser: a batch is parsed in one or two passes of <= 16 chunks each, several hundred times a second -- creating and
// joining threads per pass cost more than some passes.  Workers are created on first use and live until the process ends (detached: the
// library may be unloaded at exit while they sleep); one parse call at a time owns the pool, concurrent callers fall back to their own
// thread (the DataSpec producer thread is the only caller in practice).
class WorkerPool {
 public:
  // runs job(0 .. n-1), index 0 on the calling thread; returns when all are done
  void run(int n, const std::function<void(int)>& job) {
    if (n <= 1) { if (n == 1) job(0); return; }
    std::unique_lock<std::mutex> owner(owner_, std::try_to_lock);
    if (!owner.owns_lock()) { for (int i = 0; i < n; ++i) job(i); return; }
    {
      std::lock_guard<std::mutex> g(m_);
      while (int(workers_) < n - 1) { std::thread([this, id = workers_] { loop(id); }).detach(); ++workers_; }
      job_ = &job; n_ = n; pending_ = n - 1; ++epoch_;
    }
    wake_.notify_all();
    job(0);
    std::unique_lock<std::mutex> g(m_);
    done_.wait(g, [this] { return pending_ == 0; });
    job_ = nullptr;
  }

 private:
  void loop(size_t id) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int)>* job;
      {
        std::unique_lock<std::mutex> g(m_);
        wake_.wait(g, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (int(id) + 1 >= n_) continue;  // this pass uses fewer workers
        job = job_;
      }
      (*job)(int(id) + 1);
      std::lock_guard<std::mutex> g(m_);
      if (--pending_ == 0) done_.notify_one();
    }
  }
  std::mutex owner_, m_;
  std::condition_variable wake_, done_;
  const std::function<void(int)>* job_ = nullptr;
  size_t workers_ = 0;
  int n_ = 0, pending_ = 0;
  uint64_t epoch_ = 0;
};
// never destroyed (its workers are detached); a forked child has none of the parent's threads and starts a pool of its own
WorkerPool* g_pool = nullptr;
WorkerPool& pool() {
  static std::once_flag once;
  std::call_once(once, [] {
    g_pool = new WorkerPool();
    pthread_atfork(nullptr, nullptr, [] { g_pool = new WorkerPool(); });
  });
  return *g_pool;
}

template <class Fn>
int run_threads(int32_t B, int32_t n_threads, Fn&& fn) {  // fn(first, last, RecordError*)
  int T = std::max(1, std::min<int>(n_threads, B));
  std::vector<RecordError> errs;
  errs.resize(size_t(T));
  if (T == 1) fn(0, B, &errs[0]);
  else {
    const std::function<void(int)> job = [&](int t) {
      const int32_t lo = int32_t(int64_t(B) * t / T), hi = int32_t(int64_t(B) * (t + 1) / T);
      fn(lo, hi, &errs[size_t(t)]);
    };
    pool().run(T, job);
  }
  for (auto& e : errs)  // chunks are in record order: the first failing chunk holds the lowest failing record
    if (e.code != FDIO_OK) return fdio::fail(e.code, "%s", e.text.c_str());
  return FDIO_OK;
}

}  // namespace

extern "C" {

fdio_schema* fdio_schema_create(const fdio_column* columns, int32_t n) {
  if (!columns || n <= 0) { fdio::fail(FDIO_ERR_ARG, "fdio_schema_create: no columns"); return nullptr; }
  auto s = std::make_unique<fdio_schema>();
  for (int32_t i = 0; i < n; ++i) {
    const fdio_column& in = columns[i];
    Column c;
    if (!in.name) { fdio::fail(FDIO_ERR_ARG, "column %d has no name", i); return nullptr; }
    c.name = in.name;
    c.is_sequence = in.is_sequence != 0;
    c.dtype = in.dtype; c.width = in.width; c.transform = in.transform; c.output = in.output;
    c.num_oov = in.num_oov_indices; c.has_mask = in.has_mask != 0;
    c.mask_int = in.mask_int;
    if (in.mask_str) c.mask_str = in.mask_str;
    if (c.dtype < FDIO_INT64 || c.dtype > FDIO_STRING || c.width <= 0 || c.transform < FDIO_NONE || c.transform > FDIO_DISCRETIZE ||
        c.output < FDIO_OUT_INT32 || c.output > FDIO_OUT_SKIP) {
      fdio::fail(FDIO_ERR_ARG, "column '%s': invalid dtype / width / transform / output", in.name);
      return nullptr;
    }
    const bool yields_int = c.transform != FDIO_NONE || c.dtype == FDIO_INT64;
    const bool ok = c.output == FDIO_OUT_SKIP || (yields_int && c.output == FDIO_OUT_INT32) ||
                    (c.transform == FDIO_NONE && c.dtype == FDIO_FLOAT32 && c.output == FDIO_OUT_FLOAT32) ||
                    (c.transform == FDIO_NONE && c.dtype == FDIO_STRING && c.output == FDIO_OUT_SPAN);
    if (!ok) { fdio::fail(FDIO_ERR_ARG, "column '%s': output kind does not fit dtype and transform", in.name); return nullptr; }
    if (c.transform == FDIO_LOOKUP) {
      if (c.dtype == FDIO_FLOAT32) { fdio::fail(FDIO_ERR_ARG, "column '%s': lookup needs a string or int64 column", in.name); return nullptr; }
      if (c.num_oov < 0 || c.num_oov > 1) {
        fdio::fail(FDIO_ERR_UNSUPPORTED, "column '%s': num_oov_indices = %d (hashed OOV buckets) is not supported", in.name, c.num_oov);
        return nullptr;
      }
      c.first_vocab_index = (c.has_mask ? 1 : 0) + c.num_oov;
      for (int32_t k = 0; k < in.vocab_size; ++k) {
        bool fresh = c.dtype == FDIO_STRING ? c.str_index.insert(in.vocab_str[k], c.first_vocab_index + k)
                                            : c.int_index.emplace(in.vocab_int[k], c.first_vocab_index + k).second;
        if (!fresh) { fdio::fail(FDIO_ERR_ARG, "column '%s': repeated vocabulary term at position %d", in.name, k); return nullptr; }
      }
      c.pad_code = c.dtype == FDIO_STRING ? lookup_str(c, reinterpret_cast<const uint8_t*>(""), 0, &c.pad_i32) : lookup_int(c, 0, &c.pad_i32);
    } else if (c.transform == FDIO_DISCRETIZE) {
      if (c.dtype == FDIO_STRING || in.n_boundaries < 0 || (in.n_boundaries && !in.boundaries)) {
        fdio::fail(FDIO_ERR_ARG, "column '%s': discretize needs a numeric column and boundaries", in.name);
        return nullptr;
      }
      c.boundaries.assign(in.boundaries, in.boundaries + in.n_boundaries);
      if (!std::is_sorted(c.boundaries.begin(), c.boundaries.end())) { fdio::fail(FDIO_ERR_ARG, "column '%s': boundaries must ascend", in.name); return nullptr; }
      c.pad_i32 = bucketize(c, 0.f);
    }
    auto& index = c.is_sequence ? s->sequence_index : s->context_index;
    if (!index.emplace(c.name, int(i)).second) { fdio::fail(FDIO_ERR_ARG, "column '%s' appears twice", in.name); return nullptr; }
    s->cols.push_back(std::move(c));
  }
  return s.release();
}

void fdio_schema_destroy(fdio_schema* s) { delete s; }

int fdio_batch_steps(const fdio_schema* s, const uint8_t* const* records, const uint64_t* lens, int32_t B, int32_t* max_steps, int32_t n_threads) {
  if (!s || !records || !lens || !max_steps || B < 0) return fdio::fail(FDIO_ERR_ARG, "fdio_batch_steps: bad argument");
  return run_threads(B, n_threads, [&](int32_t lo, int32_t hi, RecordError* err) {
    std::vector<Span> found(s->cols.size());
    for (int32_t b = lo; b < hi; ++b) {
      if (split_example(records[b], lens[b], *s, &found) != FDIO_OK) {
        err->code = FDIO_ERR_CORRUPT; err->text = "record " + std::to_string(b) + ": malformed SequenceExample"; return;
      }
      int64_t m = 0;
      for (size_t ci = 0; ci < s->cols.size(); ++ci) {
        if (!s->cols[ci].is_sequence || !found[ci].present) continue;
        int64_t n = count_steps(found[ci]);
        if (n < 0) { describe(err, FDIO_ERR_CORRUPT, b, s->cols[ci], -1, "malformed protobuf"); return; }
        m = std::max(m, n);
      }
      max_steps[b] = int32_t(m);
    }
  });
}

int fdio_parse_batch(const fdio_schema* s, const uint8_t* const* records, const uint64_t* lens, int32_t B, int32_t S, void* const* out,
                     int32_t n_threads) {
  if (!s || !records || !lens || !out || B < 0 || S < 0) return fdio::fail(FDIO_ERR_ARG, "fdio_parse_batch: bad argument");
  for (size_t ci = 0; ci < s->cols.size(); ++ci)
    if (s->cols[ci].output != FDIO_OUT_SKIP && !out[ci] && B > 0 && (S > 0 || !s->cols[ci].is_sequence))
      return fdio::fail(FDIO_ERR_ARG, "fdio_parse_batch: no output buffer for column '%s'", s->cols[ci].name.c_str());
  return run_threads(B, n_threads, [&](int32_t lo, int32_t hi, RecordError* err) {
    std::vector<Span> found(s->cols.size());
    for (int32_t b = lo; b < hi && err->code == FDIO_OK; ++b) parse_record(*s, records[b], lens[b], b, S, out, &found, err);
  });
}

int fdio_parse_batch_packed(const fdio_schema* s, const uint8_t* const* records, const uint64_t* lens, int32_t B, int32_t S, void* const* out,
                            int32_t length_column, fdio_pack_column* pack, int32_t n_pack, int32_t n_threads) {
  if (!s || !records || !lens || !out || B < 0 || S < 0 || (n_pack > 0 && !pack) || n_pack < 0)
    return fdio::fail(FDIO_ERR_ARG, "fdio_parse_batch_packed: bad argument");
  const int32_t ncol = int32_t(s->cols.size());
  if (length_column < 0 || length_column >= ncol || s->cols[size_t(length_column)].is_sequence || s->cols[size_t(length_column)].output != FDIO_OUT_INT32)
    return fdio::fail(FDIO_ERR_ARG, "fdio_parse_batch_packed: length_column must be an int32 context column");
  std::vector<uint8_t> skip(size_t(ncol), 0);
  for (int32_t k = 0; k < n_pack; ++k) {
    const fdio_pack_column& pc = pack[k];
    if (pc.column < 0 || pc.column >= ncol || skip[size_t(pc.column)]) return fdio::fail(FDIO_ERR_ARG, "fdio_parse_batch_packed: bad or repeated packed column");
    const Column& c = s->cols[size_t(pc.column)];
    if (!c.is_sequence || c.transform != FDIO_NONE || c.dtype != FDIO_FLOAT32 || c.output != FDIO_OUT_FLOAT32)
      return fdio::fail(FDIO_ERR_ARG, "column '%s': only untransformed float32 sequence columns can be packed", c.name.c_str());
    if (pc.cond_column >= 0) {
      if (pc.cond_column >= ncol || !pc.cond_mask || pc.cond_n <= 0) return fdio::fail(FDIO_ERR_ARG, "column '%s': bad gating column", c.name.c_str());
      const Column& g = s->cols[size_t(pc.cond_column)];
      if (!g.is_sequence || g.width != 1 || g.output != FDIO_OUT_INT32) return fdio::fail(FDIO_ERR_ARG, "column '%s': the gating column must be an int32 sequence column of width 1", c.name.c_str());
    }
    if (!pc.rowmap && B > 0 && S > 0) return fdio::fail(FDIO_ERR_ARG, "column '%s': no row map", c.name.c_str());
    skip[size_t(pc.column)] = 1;
  }
  for (int32_t ci = 0; ci < ncol; ++ci)
    if (s->cols[size_t(ci)].output != FDIO_OUT_SKIP && !out[ci] && B > 0 && (S > 0 || !s->cols[size_t(ci)].is_sequence))
      return fdio::fail(FDIO_ERR_ARG, "fdio_parse_batch_packed: no output buffer for column '%s'", s->cols[size_t(ci)].name.c_str());
  // pass A: every other column (the length and gate columns among them)
  int code = run_threads(B, n_threads, [&](int32_t lo, int32_t hi, RecordError* err) {
    std::vector<Span> found(s->cols.size());
    for (int32_t b = lo; b < hi && err->code == FDIO_OK; ++b) parse_record(*s, records[b], lens[b], b, S, out, &found, err, skip.data());
  });
  if (code != FDIO_OK) return code;
  // rows per document and packed column -> first row of every document
  std::vector<int64_t> row0(size_t(n_pack) * size_t(B + 1), 0);
  for (int32_t k = 0; k < n_pack; ++k) {
    int64_t* r = row0.data() + size_t(k) * size_t(B + 1);
    for (int32_t b = 0; b < B; ++b) {
      int64_t n = 0;
      for (int32_t t = 0; t < S; ++t) {
        const int carry = carries(*s, pack[k], out, length_column, b, S, t);
        if (carry < 0) return fdio::fail(FDIO_ERR_ARG, "record %d, key '%s', index %d: the gating column's value is outside the loss_condition mask", b,
                                         s->cols[size_t(pack[k].column)].name.c_str(), t);
        n += carry;
      }
      r[b + 1] = r[b] + n;
    }
    pack[k].n_rows = r[B];
    if (r[B] > pack[k].capacity_rows) return fdio::fail(FDIO_ERR_ARG, "column '%s': %lld rows in use, capacity %lld", s->cols[size_t(pack[k].column)].name.c_str(),
                                                        (long long)r[B], (long long)pack[k].capacity_rows);
  }
  // pass B: the packed columns, every document at its own rows
  return run_threads(B, n_threads, [&](int32_t lo, int32_t hi, RecordError* err) {
    std::vector<Span> found(s->cols.size());
    std::vector<float> scratch_row;
    for (int32_t b = lo; b < hi && err->code == FDIO_OK; ++b) {
      if (split_example(records[b], lens[b], *s, &found) != FDIO_OK) { err->code = FDIO_ERR_CORRUPT; err->text = "record " + std::to_string(b) + ": malformed SequenceExample"; return; }
      for (int32_t k = 0; k < n_pack && err->code == FDIO_OK; ++k) {
        scratch_row.resize(size_t(s->cols[size_t(pack[k].column)].width));
        pack_record(*s, pack[k], found[size_t(pack[k].column)], b, S, out, length_column, row0[size_t(k) * size_t(B + 1) + size_t(b)], scratch_row.data(), err);
      }
    }
  });
}

}  // extern "C"
