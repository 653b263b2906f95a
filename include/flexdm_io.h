/* flexdm_io.h -- C ABI of the data formats either side of the MFP step.
 *
 *   input side : TFRecord files of tf.train.SequenceExample -> padded batch columns, what
 *                DataSpec.make_dataset / DataSpec.parse_fn produce (src/mfp/mfp/data/spec.py:213-287,
 *                data/discretizer.py:6-31);
 *   weight side: TensorFlow tensor-bundle checkpoints (best.ckpt / final.ckpt = <prefix>.index +
 *                <prefix>.data-00000-of-00001), what Model.save_weights / load_weights read and write
 *                (train.py:67-69,94-97; eval.py:169-172; helpers/callbacks.py:49-56).
 *
 * The reference does both through TensorFlow's C++ runtime (tf.data.TFRecordDataset,
 * tf.io.parse_sequence_example, Keras StringLookup / IntegerLookup / Discretization, tf.train.Checkpoint);
 * TensorFlow is not part of /root/reference, so each entry point states the published format it follows.
 * flex_dm_b200/dataspec.py and flex_dm_b200/checkpoint.py are the ctypes bindings.
 *
 * Conventions: host pointers only (this library never touches the GPU; batch outputs are written straight
 * into caller-owned buffers, normally pinned memory that DevicePrefetcher copies from).  Calls return
 * FDIO_OK (0) or a negative code; fdio_last_error() gives the message (thread-local).  Handles are
 * immutable after creation and may be shared between threads.
 */
#ifndef FLEXDM_IO_H_
#define FLEXDM_IO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { FDIO_OK = 0, FDIO_ERR_ARG = -1, FDIO_ERR_IO = -2, FDIO_ERR_CORRUPT = -3, FDIO_ERR_INVALID = -4, FDIO_ERR_OOV = -5,
       FDIO_ERR_UNSUPPORTED = -6, FDIO_ERR_NOT_FOUND = -7 };

const char* fdio_last_error(void);
int fdio_version(void);

/* ---- CRC-32C (Castagnoli), the checksum of both formats -------------------------------------------- */
uint32_t fdio_crc32c(const void* data, size_t n);
/* continue a running CRC: fdio_crc32c_extend(fdio_crc32c(a), b) == fdio_crc32c(a ++ b) */
uint32_t fdio_crc32c_extend(uint32_t crc, const void* data, size_t n);
/* TFRecord / tensor-bundle "masked" form: rotr(crc, 15) + 0xa282ead8 */
uint32_t fdio_crc32c_mask(uint32_t crc);
uint32_t fdio_crc32c_unmask(uint32_t masked);

/* ---- TFRecord files (tf.data.TFRecordDataset, spec.py:231-236) --------------------------------------
 * record = uint64 length | uint32 masked crc32c(length) | bytes[length] | uint32 masked crc32c(bytes), little endian. */
typedef struct fdio_tfrecord fdio_tfrecord;
/* mmaps the file and indexes its records; verify_crc: 0 = none, 1 = length CRCs, 2 = length and payload CRCs. */
fdio_tfrecord* fdio_tfrecord_open(const char* path, int verify_crc);
void fdio_tfrecord_close(fdio_tfrecord* f);
int64_t fdio_tfrecord_count(const fdio_tfrecord* f);
/* pointers into the mapping: valid until fdio_tfrecord_close */
int fdio_tfrecord_get(const fdio_tfrecord* f, int64_t index, const uint8_t** data, uint64_t* len);
/* writes records in the same framing (used to export synthetic datasets and by the tests) */
int fdio_tfrecord_write(const char* path, const uint8_t* const* records, const uint64_t* lens, int64_t n);

/* ---- SequenceExample -> batch columns (DataSpec.parse_fn, spec.py:255-287) --------------------------- */
enum { FDIO_INT64 = 0, FDIO_FLOAT32 = 1, FDIO_STRING = 2 };                 /* column["dtype"] of the YAML spec */
enum { FDIO_NONE = 0, FDIO_LOOKUP = 1, FDIO_DISCRETIZE = 2 };               /* DataSpec._init_preprocessor, spec.py:90-105 */
enum { FDIO_OUT_INT32 = 0, FDIO_OUT_FLOAT32 = 1, FDIO_OUT_SPAN = 2, FDIO_OUT_SKIP = 3 };

/* One column of the YAML spec (data/crello-spec.yml, data/rico-spec.yml) with its preprocessor resolved.
 *  lookup     : Keras StringLookup / IntegerLookup(vocabulary, num_oov_indices, mask_token) in "int" output mode
 *               (spec.py:107-134): index = [mask_token] + [OOV] * num_oov_indices + vocabulary; a value outside the
 *               vocabulary maps to the OOV index, or fails with FDIO_ERR_OOV when num_oov_indices == 0.
 *               num_oov_indices > 1 (hashed OOV buckets) is FDIO_ERR_UNSUPPORTED.
 *  discretize : SequenceDiscretizer (discretizer.py:6-31): cast to float32, Bucketize over float32 boundaries
 *               = number of boundaries <= x.
 *  output     : int64 results are cast to int32 (spec.py:281-285); FDIO_OUT_SPAN writes (offset, length) int64 pairs
 *               addressing raw bytes inside the record (demo-only string columns id / uuid); FDIO_OUT_SKIP parses and
 *               validates the column but writes nothing. */
typedef struct {
  const char* name;
  int32_t is_sequence;          /* FixedLenSequenceFeature (1) or FixedLenFeature (0), spec.py:258-274 */
  int32_t dtype;                /* FDIO_INT64 / FDIO_FLOAT32 / FDIO_STRING */
  int32_t width;                /* prod(column.get("shape", (1,))) */
  int32_t transform;            /* FDIO_NONE / FDIO_LOOKUP / FDIO_DISCRETIZE */
  int32_t output;               /* FDIO_OUT_* */
  int32_t vocab_size;
  const char* const* vocab_str; /* string lookup: vocab_size NUL-terminated tokens */
  const int64_t* vocab_int;     /* integer lookup: vocab_size values */
  int32_t num_oov_indices;
  int32_t has_mask;             /* mask_token is not None */
  const char* mask_str;
  int64_t mask_int;
  int32_t n_boundaries;
  const float* boundaries;      /* ascending, float32 (the Bucketize attribute type) */
} fdio_column;

typedef struct fdio_schema fdio_schema;
fdio_schema* fdio_schema_create(const fdio_column* columns, int32_t n);  /* copies everything it needs */
void fdio_schema_destroy(fdio_schema* s);

/* Pass 1: the number of steps of every sequence column in each record (parse_sequence_example pads every
 * sequence feature to the longest in the batch).  max_steps[B] receives, per record, the largest step count. */
int fdio_batch_steps(const fdio_schema* s, const uint8_t* const* records, const uint64_t* lens, int32_t B, int32_t* max_steps,
                     int32_t n_threads);
/* Pass 2: fill the batch.  out[c] is the buffer of schema column c (ignored for FDIO_OUT_SKIP):
 *   context column   [B, width]      sequence column   [B, S, width]
 * of int32 / float32 / int64 pairs as the column's output kind says; steps beyond a record's own length are the
 * parse defaults (0, 0.0, "") put through the column's preprocessor, exactly like the reference's padded batch.
 * S must be >= every record's step count (S > batch max gives fixed-shape batches).  Records are parsed on
 * n_threads host threads (<= 1: the calling thread). */
int fdio_parse_batch(const fdio_schema* s, const uint8_t* const* records, const uint64_t* lens, int32_t B, int32_t S, void* const* out,
                     int32_t n_threads);

/* Pass 2, packed form (the input pipeline's batch format, flex_dm_b200.data.pack_batch / mfp_set_packed_rows): the float32 sequence
 * columns named in `pack` are written as [n_rows, width] -- only the rows of the elements that carry the column, document by document --
 * next to an int32 [B, S] element -> row map (-1 = none), instead of dense [B, S, width]; every other column as fdio_parse_batch writes
 * it.  An element (b, t) carries a packed column iff t <= the value the context column `length_column` holds for b (the zero-based
 * length DataSpec's IntegerLookup yields: valid positions, masking.py:24-53 / mask.py:21-33) and, when the column has a gate, iff
 * cond_mask[v] != 0 for the value v the width-1 int32 sequence column `cond_column` holds at (b, t) (loss_condition,
 * data/crello-spec.yml:88-121; the mask is indexed by the preprocessed value).  Everywhere else filter_padding overwrites the value
 * with <UNUSED> before the model reads it, so nothing is lost; 59 % of crello's embedding rows are never written.  out[column] must
 * hold capacity_rows rows; n_rows receives the rows in use.  Values of elements that carry nothing are still decoded (errors do not
 * depend on the format). */
typedef struct {
  int32_t column;          /* schema index: float32 sequence column, no transform, FDIO_OUT_FLOAT32 */
  int32_t cond_column;     /* schema index of the gating int32 sequence column (width 1), or -1 */
  const uint8_t* cond_mask;
  int32_t cond_n;
  int32_t* rowmap;         /* [B, S] */
  int64_t capacity_rows;
  int64_t n_rows;          /* out */
} fdio_pack_column;
int fdio_parse_batch_packed(const fdio_schema* s, const uint8_t* const* records, const uint64_t* lens, int32_t B, int32_t S, void* const* out,
                            int32_t length_column, fdio_pack_column* pack, int32_t n_pack, int32_t n_threads);

/* ---- TensorFlow tensor-bundle checkpoints (Model.load_weights / save_weights, train.py:67-69,94-97) ---
 * <prefix>.index is an immutable sorted string table (LevelDB table format: prefix-compressed blocks, restart
 * array, 5-byte block trailer, 48-byte footer with magic 0xdb4775248b80fb57); key "" holds BundleHeaderProto,
 * every other key a BundleEntryProto {dtype, shape, shard_id, offset, size, crc32c}; tensor bytes sit in
 * <prefix>.data-<shard>-of-<num_shards>.  Keras object checkpoints name variables
 * "<attribute path>/.ATTRIBUTES/VARIABLE_VALUE" and store the object graph under "_CHECKPOINTABLE_OBJECT_GRAPH". */
typedef struct fdio_bundle fdio_bundle;
fdio_bundle* fdio_bundle_open(const char* prefix);
void fdio_bundle_close(fdio_bundle* b);
int32_t fdio_bundle_count(const fdio_bundle* b);                 /* entries, header excluded */
const char* fdio_bundle_key(const fdio_bundle* b, int32_t i);    /* sorted order */
int32_t fdio_bundle_find(const fdio_bundle* b, const char* key); /* index or FDIO_ERR_NOT_FOUND */
/* dtype = TensorFlow DataType enum (DT_FLOAT = 1, DT_INT64 = 9, DT_STRING = 7 ...); dims receives up to max_rank sizes */
int fdio_bundle_info(const fdio_bundle* b, int32_t i, int32_t* dtype, int32_t* rank, int64_t* dims, int32_t max_rank, int64_t* nbytes);
/* copies the tensor bytes (verifying the entry's crc32c); DT_STRING tensors come back as their raw encoding */
int fdio_bundle_read(const fdio_bundle* b, int32_t i, void* dst, int64_t nbytes);

/* Writer: entries must be added in any order; finish sorts keys, writes <prefix>.data-00000-of-00001 and <prefix>.index. */
typedef struct fdio_bundle_writer fdio_bundle_writer;
fdio_bundle_writer* fdio_bundle_writer_create(const char* prefix);
int fdio_bundle_writer_add(fdio_bundle_writer* w, const char* key, int32_t dtype, int32_t rank, const int64_t* dims, const void* data,
                           int64_t nbytes);
int fdio_bundle_writer_finish(fdio_bundle_writer* w);            /* also frees the writer */

#ifdef __cplusplus
}
#endif
#endif /* FLEXDM_IO_H_ */
