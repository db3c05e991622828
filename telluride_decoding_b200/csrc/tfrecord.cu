// TFRecord / tf.train.Example reader (host code) for libtdgram.so.
//
// Replaces, for the drop-in boundary, `tf.data.TFRecordDataset` + `tf.io.parse_example` as the
// reference uses them to get one frame per record (telluride_decoding/brain_data.py:733-745,
// parse_and_select_from_tfrecord :758-812, discover_feature_shapes :853-892, count_tfrecords
// :895-920; the writer's wire format is ingest.py:1118-1172).  The file is parsed straight into a
// caller-owned [frames, width] float32 buffer (pinned host memory in the Python layer), which is
// then one H2D copy away from the lagged kernels.
//
// Container: every record is  uint64 length | uint32 masked_crc32c(length) | bytes | uint32
// masked_crc32c(bytes).  Payload: Example{1: Features{1: map<string, Feature>}} with
// Feature{1: BytesList, 2: FloatList{1: packed floats}, 3: Int64List{1: packed varints}}.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace td {
namespace {

// ---- CRC32C (Castagnoli), table driven --------------------------------------------------------
struct Crc32cTable {
  uint32_t t[8][256];
  Crc32cTable() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
  }
};

uint32_t crc32c(const uint8_t* p, size_t n) {
  static const Crc32cTable tab;
  uint32_t c = 0xFFFFFFFFu;
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;
    c = tab.t[7][w & 0xff] ^ tab.t[6][(w >> 8) & 0xff] ^ tab.t[5][(w >> 16) & 0xff] ^
        tab.t[4][(w >> 24) & 0xff] ^ tab.t[3][(w >> 32) & 0xff] ^ tab.t[2][(w >> 40) & 0xff] ^
        tab.t[1][(w >> 48) & 0xff] ^ tab.t[0][(w >> 56) & 0xff];
    p += 8;
    n -= 8;
  }
  while (n--) c = (c >> 8) ^ tab.t[0][(c ^ *p++) & 0xff];
  return c ^ 0xFFFFFFFFu;
}

inline uint32_t masked_crc(const uint8_t* p, size_t n) {
  const uint32_t c = crc32c(p, n);
  return ((c >> 15) | (c << 17)) + 0xa282ead8u;
}

// ---- protobuf wire helpers -----------------------------------------------------------------------
struct Span {
  const uint8_t* p;
  const uint8_t* end;
  bool empty() const { return p >= end; }
};

bool read_varint(Span& s, uint64_t* v) {
  uint64_t r = 0;
  for (int shift = 0; shift < 64 && s.p < s.end; shift += 7) {
    const uint8_t b = *s.p++;
    r |= (uint64_t)(b & 0x7f) << shift;
    if (!(b & 0x80)) { *v = r; return true; }
  }
  return false;
}

// Reads one field header; for length-delimited fields `sub` is the payload, for the others the
// value is skipped (fixed32 payloads are exposed through `sub` as 4 bytes).
bool read_field(Span& s, uint32_t* field, uint32_t* wire, Span* sub) {
  uint64_t key;
  if (!read_varint(s, &key)) return false;
  *field = (uint32_t)(key >> 3);
  *wire = (uint32_t)(key & 7);
  uint64_t len;
  switch (*wire) {
    case 0: return read_varint(s, &len);
    case 1: if (s.end - s.p < 8) return false; sub->p = s.p; sub->end = s.p + 8; s.p += 8; return true;
    case 2:
      if (!read_varint(s, &len) || (uint64_t)(s.end - s.p) < len) return false;
      sub->p = s.p; sub->end = s.p + len; s.p += len; return true;
    case 5: if (s.end - s.p < 4) return false; sub->p = s.p; sub->end = s.p + 4; s.p += 4; return true;
    default: return false;
  }
}

enum FeatureKind { KIND_NONE = 0, KIND_BYTES = 1, KIND_FLOAT = 2, KIND_INT64 = 3 };

// Visits the (name, Feature payload) pairs of one serialized tf.train.Example.
template <typename Fn>
bool for_each_feature(Span example, Fn&& fn) {
  while (!example.empty()) {
    uint32_t f, w;
    Span features{nullptr, nullptr};
    if (!read_field(example, &f, &w, &features)) return false;
    if (f != 1 || w != 2) continue;                       // Example.features
    while (!features.empty()) {
      Span entry{nullptr, nullptr};
      if (!read_field(features, &f, &w, &entry)) return false;
      if (f != 1 || w != 2) continue;                     // Features.feature map entry
      Span key{nullptr, nullptr}, value{nullptr, nullptr};
      while (!entry.empty()) {
        Span sub{nullptr, nullptr};
        if (!read_field(entry, &f, &w, &sub)) return false;
        if (w != 2) continue;
        if (f == 1) key = sub; else if (f == 2) value = sub;
      }
      if (key.p && !fn(key, value)) return false;
    }
  }
  return true;
}

// Counts (out == nullptr) or stores the values of a Feature as float32.  Returns the kind.
int feature_values(Span feature, float* out, int64_t cap, int64_t* count) {
  int kind = KIND_NONE;
  int64_t n = 0;
  while (!feature.empty()) {
    uint32_t f, w;
    Span list{nullptr, nullptr};
    if (!read_field(feature, &f, &w, &list)) return -1;
    if (w != 2) continue;
    if (f == 1) { kind = kind ? kind : KIND_BYTES; continue; }
    if (f != 2 && f != 3) continue;
    kind = (f == 2) ? KIND_FLOAT : KIND_INT64;
    while (!list.empty()) {
      uint32_t lf, lw;
      Span vals{nullptr, nullptr};
      if (f == 2) {
        if (!read_field(list, &lf, &lw, &vals)) return -1;
        if (lf != 1 || (lw != 2 && lw != 5)) continue;    // packed (usual) or one fixed32 per value
        const int64_t m = (vals.end - vals.p) / 4;
        if (out) {
          if (n + m > cap) return -2;
          memcpy(out + n, vals.p, (size_t)m * 4);
        }
        n += m;
      } else {
        uint64_t key;
        Span save = list;
        if (!read_varint(list, &key)) return -1;
        if ((key >> 3) != 1) { list = save; if (!read_field(list, &lf, &lw, &vals)) return -1; continue; }
        if ((key & 7) == 2) {                              // packed varints
          uint64_t len;
          if (!read_varint(list, &len) || (uint64_t)(list.end - list.p) < len) return -1;
          Span pv{list.p, list.p + len};
          list.p += len;
          while (!pv.empty()) {
            uint64_t v;
            if (!read_varint(pv, &v)) return -1;
            if (out) { if (n >= cap) return -2; out[n] = (float)(int64_t)v; }
            ++n;
          }
        } else if ((key & 7) == 0) {
          uint64_t v;
          if (!read_varint(list, &v)) return -1;
          if (out) { if (n >= cap) return -2; out[n] = (float)(int64_t)v; }
          ++n;
        } else {
          return -1;
        }
      }
    }
  }
  *count = n;
  return kind;
}

// Walks the container.  fn(payload span, record index) -> false stops with an error already set.
template <typename Fn>
int for_each_record(const uint8_t* bytes, size_t n_bytes, int verify_crc, int64_t max_records,
                    int64_t* n_records, Fn&& fn) {
  size_t off = 0;
  int64_t n = 0;
  *n_records = 0;
  while (off < n_bytes && (max_records < 0 || n < max_records)) {
    *n_records = n;                                   // good records so far, also on an error return
    if (n_bytes - off < 12) {
      set_error("td_tfrecord: truncated record header at byte %zu (record %lld)", off, (long long)n);
      return TD_ERR_INVALID;
    }
    uint64_t len;
    uint32_t len_crc;
    memcpy(&len, bytes + off, 8);
    memcpy(&len_crc, bytes + off + 8, 4);
    if (verify_crc >= 1 && masked_crc(bytes + off, 8) != len_crc) {
      set_error("td_tfrecord: corrupted length checksum at byte %zu (record %lld)", off, (long long)n);
      return TD_ERR_INVALID;
    }
    if (len > n_bytes - off - 12 || n_bytes - off - 12 - len < 4) {
      set_error("td_tfrecord: truncated record %lld (%llu payload bytes at byte %zu of %zu)",
                (long long)n, (unsigned long long)len, off, n_bytes);
      return TD_ERR_INVALID;
    }
    const uint8_t* payload = bytes + off + 12;
    if (verify_crc >= 2) {
      uint32_t data_crc;
      memcpy(&data_crc, payload + len, 4);
      if (masked_crc(payload, (size_t)len) != data_crc) {
        set_error("td_tfrecord: corrupted data checksum in record %lld", (long long)n);
        return TD_ERR_INVALID;
      }
    }
    if (!fn(Span{payload, payload + len}, n)) return TD_ERR_INVALID;
    off += 12 + (size_t)len + 4;
    ++n;
  }
  *n_records = n;
  return TD_OK;
}

}  // namespace
}  // namespace td

using namespace td;

extern "C" int td_tfrecord_count(const void* file_bytes, size_t n_bytes, int verify_crc,
                                 int64_t* n_records) {
  TD_REQUIRE(file_bytes && n_records, "td_tfrecord_count: null argument");
  return for_each_record(static_cast<const uint8_t*>(file_bytes), n_bytes, verify_crc, -1, n_records,
                         [](Span, int64_t) { return true; });
}

extern "C" int td_tfrecord_features(const void* file_bytes, size_t n_bytes, char* names,
                                    size_t names_cap, int32_t* widths, int32_t* kinds,
                                    int32_t max_features, int32_t* n_features) {
  TD_REQUIRE(file_bytes && names && widths && kinds && n_features, "td_tfrecord_features: null argument");
  int32_t nf = 0;
  size_t used = 0;
  bool overflow = false;
  int64_t n_rec = 0;
  int rc = for_each_record(
      static_cast<const uint8_t*>(file_bytes), n_bytes, 1, 1, &n_rec, [&](Span payload, int64_t) {
        const bool ok = for_each_feature(payload, [&](Span key, Span value) {
          int64_t count = 0;
          const int kind = feature_values(value, nullptr, 0, &count);
          if (kind < 0) return false;
          const size_t klen = (size_t)(key.end - key.p);
          if (nf >= max_features || used + klen + 1 > names_cap) { overflow = true; return true; }
          memcpy(names + used, key.p, klen);
          names[used + klen] = '\0';
          used += klen + 1;
          widths[nf] = (int32_t)count;
          kinds[nf] = kind;
          ++nf;
          return true;
        });
        if (!ok) set_error("td_tfrecord_features: the first record is not a tf.train.Example");
        return ok;
      });
  if (rc) return rc;
  TD_REQUIRE(n_rec == 1, "td_tfrecord_features: the file holds no record");
  TD_REQUIRE(!overflow, "td_tfrecord_features: more than %d features or names longer than %zu bytes",
             max_features, names_cap);
  *n_features = nf;
  return TD_OK;
}

extern "C" int td_tfrecord_read_f32(const void* file_bytes, size_t n_bytes, const char* const* names,
                                    const int32_t* widths, int32_t n_names, float* out, int64_t ld_out,
                                    int64_t max_records, int verify_crc, int64_t* n_records) {
  TD_REQUIRE(file_bytes && names && widths && out && n_records && n_names > 0,
             "td_tfrecord_read_f32: null argument");
  std::vector<int64_t> col0(n_names);
  std::vector<size_t> name_len(n_names);
  int64_t total = 0;
  for (int i = 0; i < n_names; ++i) {
    TD_REQUIRE(names[i] && widths[i] > 0, "td_tfrecord_read_f32: bad feature %d", i);
    col0[i] = total;
    total += widths[i];
    name_len[i] = strlen(names[i]);
  }
  TD_REQUIRE(ld_out >= total, "td_tfrecord_read_f32: ld_out (%lld) < total width (%lld)",
             (long long)ld_out, (long long)total);
  std::vector<char> seen(n_names);
  return for_each_record(
      static_cast<const uint8_t*>(file_bytes), n_bytes, verify_crc, max_records, n_records,
      [&](Span payload, int64_t r) {
        std::fill(seen.begin(), seen.end(), 0);
        float* row = out + r * ld_out;
        bool bad_width = false;
        int bad_i = -1;
        int64_t bad_count = 0;
        const bool ok = for_each_feature(payload, [&](Span key, Span value) {
          const size_t klen = (size_t)(key.end - key.p);
          for (int i = 0; i < n_names; ++i) {
            if (klen != name_len[i] || memcmp(key.p, names[i], klen) != 0) continue;
            int64_t count = 0;
            const int kind = feature_values(value, row + col0[i], widths[i], &count);
            if (kind == -1) return false;
            if (kind == -2 || count != widths[i]) { bad_width = true; bad_i = i; bad_count = count; return true; }
            seen[i] = 1;
          }
          return true;
        });
        if (!ok) { set_error("td_tfrecord_read_f32: record %lld is not a tf.train.Example", (long long)r); return false; }
        if (bad_width) {
          set_error("td_tfrecord_read_f32: feature %s of record %lld does not have %d values (%lld found)",
                    names[bad_i], (long long)r, widths[bad_i], (long long)bad_count);
          return false;
        }
        for (int i = 0; i < n_names; ++i)
          if (!seen[i]) {
            set_error("td_tfrecord_read_f32: Could not find all desired features (%s) in record %lld",
                      names[i], (long long)r);
            return false;
          }
        return true;
      });
}
