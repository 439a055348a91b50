// table.cu — context (stream, pool, profiling) and Vec<RecordBatch> <-> HBM through the Arrow C Data Interface.
//
// HBM layout: the batches of one Vec<RecordBatch> are concatenated column-wise.  A column is one contiguous
// device allocation of its values (8/4-byte little endian, or bit-packed booleans, or Utf8 int32 offsets +
// bytes) plus an optional Arrow validity bitmap; `batch_offsets` remembers where each input batch started so
// operators that the reference runs per batch (Filter) can hand back the same batch structure.
// Allocations are padded by >= 256 B so vector loads may over-read the tail safely.
#include <cstdlib>
#include <cstring>

#include "qgpu_internal.h"

namespace qgpu {

// ---------------------------------------------------------------- Ctx
// Operators allocate and free multi-hundred-MB temporaries on every call.  cudaMallocAsync alone fragments its pool
// under that pattern (a 60 MB request carved out of a freed 480 MB block forces the next 480 MB request to map new
// physical memory: ~1 ms of host time with the GPU idle), so freed blocks are kept in exact-size free lists and handed
// back to the next request of the same rounded size.  Everything a context does is ordered on its single stream,
// so a block may be reused the moment the host frees it.
static size_t round_block(size_t bytes) {
  if (bytes == 0) bytes = 256;
  if (bytes < (1u << 20)) return (bytes + 511) & ~(size_t)511;
  return (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
}
void* Ctx::alloc(size_t bytes) {
  const size_t sz = round_block(bytes);
  std::lock_guard<std::mutex> lk(cache_mu);
  auto it = free_blocks.find(sz);
  if (it != free_blocks.end() && !it->second.empty()) {
    void* p = it->second.back();
    it->second.pop_back();
    cached_bytes -= sz;
    live_blocks[p] = sz;
    return p;
  }
  void* p = nullptr;
  cudaError_t e = cudaMallocAsync(&p, sz, stream);
  if (e == cudaErrorMemoryAllocation) {  // give the cache back and retry once
    cudaGetLastError();
    trim_locked(0);
    e = cudaMallocAsync(&p, sz, stream);
  }
  QG_CUDA(e);
  live_blocks[p] = sz;
  return p;
}
void Ctx::free(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(cache_mu);
  auto it = live_blocks.find(p);
  if (it == live_blocks.end()) {
    cudaFreeAsync(p, stream);
    return;
  }
  const size_t sz = it->second;
  live_blocks.erase(it);
  free_blocks[sz].push_back(p);
  cached_bytes += sz;
  const size_t cap = (size_t)opt("pool.max_cached_bytes", (int64_t)64 << 30);
  if (cached_bytes > cap) trim_locked(cap / 2);
}
void Ctx::trim(size_t keep_bytes) {
  std::lock_guard<std::mutex> lk(cache_mu);
  trim_locked(keep_bytes);
}
void Ctx::trim_locked(size_t keep_bytes) {
  for (auto it = free_blocks.rbegin(); it != free_blocks.rend() && cached_bytes > keep_bytes; ++it) {
    while (!it->second.empty() && cached_bytes > keep_bytes) {
      cudaFreeAsync(it->second.back(), stream);
      it->second.pop_back();
      cached_bytes -= it->first;
    }
  }
}
void* Ctx::zalloc(size_t bytes) {
  void* p = alloc(bytes);
  QG_CUDA(cudaMemsetAsync(p, 0, bytes == 0 ? 256 : bytes, stream));
  return p;
}
void Ctx::prof_begin(const char* name) {
  ++launch_count;
  if (!profile_events) {
    return;
  }
  ProfileEntry e;
  e.name = name;
  e.timed = true;
  QG_CUDA(cudaEventCreate(&e.start));
  QG_CUDA(cudaEventCreate(&e.stop));
  QG_CUDA(cudaEventRecord(e.start, stream));
  prof.push_back(e);
}
void Ctx::prof_end() {
  if (!profile_events || prof.empty()) return;
  cudaEventRecord(prof.back().stop, stream);
}
void Ctx::sync() { QG_CUDA(cudaStreamSynchronize(stream)); }
void* Ctx::scratch(size_t bytes) {
  if (bytes > pinned_scratch_bytes) {
    if (pinned_scratch) {
      QG_CUDA(cudaStreamSynchronize(stream));
      cudaFreeHost(pinned_scratch);
    }
    size_t nb = bytes < 65536 ? 65536 : bytes * 2;
    QG_CUDA(cudaMallocHost(&pinned_scratch, nb));
    pinned_scratch_bytes = nb;
  }
  return pinned_scratch;
}

int resident_ctas_per_sm(const void* kernel, int threads, size_t smem) {
  static std::mutex mu;
  static std::map<std::pair<const void*, size_t>, int> cache;
  std::lock_guard<std::mutex> lk(mu);
  auto key = std::make_pair(kernel, smem * 4096 + (size_t)threads);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  int per_sm = 0;
  QG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
  if (per_sm < 1) per_sm = 1;
  cache[key] = per_sm;
  return per_sm;
}

void Table::adopt(Ctx* c) {
  ctx = c;
  if (c) {
    std::lock_guard<std::mutex> lk(c->cache_mu);
    c->live_tables.insert(this);
  }
}

void drain_all_lanes(Ctx* c);  // capi.cu

Table::~Table() {
  if (ready) cudaEventDestroy(ready);
  if (!ctx) return;  // schema-only, or the context is gone (its memory went with it)
  // another lane's stream may still be reading these buffers: the owner's cache hands blocks out again in ITS stream
  // order only, so let every lane finish first
  if (shared) drain_all_lanes(ctx);
  {
    std::lock_guard<std::mutex> lk(ctx->cache_mu);
    ctx->live_tables.erase(this);
  }
  for (auto& c : cols) free_column(ctx, c);
}

Column alloc_column(Ctx* c, uint8_t dtype, int64_t n, bool with_valid) {
  Column col;
  col.dtype = dtype;
  col.n = n;
  col.owned = true;
  switch (dtype) {
    case DT_I32: case DT_F32: col.data = c->alloc(pad_bytes((size_t)n * 4)); break;
    case DT_I64: case DT_F64: col.data = c->alloc(pad_bytes((size_t)n * 8)); break;
    case DT_BOOL: col.data = c->zalloc(bitmap_bytes(n)); break;
    case DT_UTF8: col.data = c->alloc(pad_bytes((size_t)(n + 1) * 4)); break;
    default: break;
  }
  if (with_valid) col.valid = (uint32_t*)c->zalloc(bitmap_bytes(n));
  return col;
}

void free_column(Ctx* c, Column& col) {
  if (!col.owned) {
    if (col.owned_valid) {
      c->free(col.valid);
      col.valid = nullptr;
      col.owned_valid = false;
    }
    return;
  }
  c->free(col.data);
  c->free(col.valid);
  c->free(col.chars);
  col.data = nullptr;
  col.valid = nullptr;
  col.chars = nullptr;
}

SchemaView schema_view(const Table& t) {
  SchemaView s;
  for (size_t i = 0; i < t.cols.size(); ++i) {
    s.dtypes.push_back(t.cols[i].dtype);
    // a schema-only table (ctx == nullptr) has no buffers to inspect: trust the field's nullable flag
    s.has_nulls.push_back(t.cols[i].valid != nullptr || t.cols[i].dtype == DT_NULL || (t.ctx == nullptr && t.fields[i].nullable));
  }
  return s;
}

// ---------------------------------------------------------------- bit utilities (host)
static inline bool get_bit(const uint8_t* b, int64_t i) { return (b[i >> 3] >> (i & 7)) & 1; }

static void copy_bits(uint8_t* dst, int64_t dst_off, const uint8_t* src, int64_t src_off, int64_t n) {
  if (n <= 0) return;
  if ((dst_off & 7) == 0 && (src_off & 7) == 0) {
    memcpy(dst + (dst_off >> 3), src + (src_off >> 3), (size_t)(n >> 3));
    for (int64_t i = n & ~(int64_t)7; i < n; ++i)
      if (get_bit(src, src_off + i)) dst[(dst_off + i) >> 3] |= (uint8_t)(1u << ((dst_off + i) & 7));
    return;
  }
  for (int64_t i = 0; i < n; ++i)
    if (get_bit(src, src_off + i)) dst[(dst_off + i) >> 3] |= (uint8_t)(1u << ((dst_off + i) & 7));
}

static int64_t count_bits(const uint8_t* b, int64_t off, int64_t n) {
  int64_t c = 0;
  for (int64_t i = 0; i < n; ++i) c += get_bit(b, off + i);
  return c;
}

static uint8_t dtype_of_format(const char* f) {
  if (!f) return DT_OTHER;
  if (!strcmp(f, "b")) return DT_BOOL;
  if (!strcmp(f, "i")) return DT_I32;
  if (!strcmp(f, "l")) return DT_I64;
  if (!strcmp(f, "f")) return DT_F32;
  if (!strcmp(f, "g")) return DT_F64;
  if (!strcmp(f, "u")) return DT_UTF8;
  if (!strcmp(f, "n")) return DT_NULL;
  return DT_OTHER;
}
static const char* format_of_dtype(uint8_t d) {
  switch (d) {
    case DT_BOOL: return "b";
    case DT_I32: return "i";
    case DT_I64: return "l";
    case DT_F32: return "f";
    case DT_F64: return "g";
    case DT_UTF8: return "u";
    default: return "n";
  }
}

// ---------------------------------------------------------------- import
static void read_fields(const ArrowSchema* schema, Table& t) {
  if (!schema || !schema->format || strcmp(schema->format, "+s") != 0)
    throw_error(QGPU_ERR_INVALID, "expected a struct schema (the schema of a RecordBatch)");
  for (int64_t i = 0; i < schema->n_children; ++i) {
    const ArrowSchema* cs = schema->children[i];
    Field f;
    f.name = cs->name ? cs->name : "";
    f.dtype = dtype_of_format(cs->format);
    f.nullable = (cs->flags & ARROW_FLAG_NULLABLE) != 0;
    if (f.dtype == DT_OTHER)
      throw_error(QGPU_ERR_UNSUPPORTED, "column '%s' has Arrow format '%s', which is not on the accelerated path", f.name.c_str(),
                  cs->format ? cs->format : "?");
    t.fields.push_back(f);
  }
}

std::shared_ptr<Table> table_from_schema(ArrowSchema* schema) {
  auto t = std::make_shared<Table>();
  t->ctx = nullptr;  // owns no device memory
  read_fields(schema, *t);
  t->batch_offsets.push_back(0);
  t->cols.resize(t->fields.size());
  for (size_t i = 0; i < t->fields.size(); ++i) t->cols[i].dtype = t->fields[i].dtype;
  if (schema->release) schema->release(schema);
  return t;
}

// Ownership of the inputs: the C structs are released before return on success AND on failure (qgpu_table_import
// does it, capi.cu), so the caller never has to guess; the buffers are only read during the call.
std::shared_ptr<Table> import_batches(Ctx* c, ArrowArray* batches, int64_t n_batches, ArrowSchema* schema) {
  if (!schema || !schema->format || strcmp(schema->format, "+s") != 0)
    throw_error(QGPU_ERR_INVALID, "qgpu_table_import expects a struct schema (one RecordBatch per ArrowArray)");
  auto t = std::make_shared<Table>();
  t->adopt(c);
  const int64_t ncol = schema->n_children;
  read_fields(schema, *t);
  t->batch_offsets.push_back(0);
  int64_t total = 0;
  for (int64_t b = 0; b < n_batches; ++b) {
    if (batches[b].n_children != ncol) throw_error(QGPU_ERR_INVALID, "batch %lld has %lld columns, schema has %lld", (long long)b,
                                                    (long long)batches[b].n_children, (long long)ncol);
    for (int64_t ci = 0; ci < ncol; ++ci) {
      const ArrowArray* a = batches[b].children[ci];
      if (!a || a->length < batches[b].offset + batches[b].length)
        throw_error(QGPU_ERR_INVALID, "batch %lld column %lld is shorter than the struct array's offset + length", (long long)b,
                    (long long)ci);
    }
    total += batches[b].length;
    t->batch_offsets.push_back(total);
  }
  if (total >= 0xfffffff0ll) throw_error(QGPU_ERR_UNSUPPORTED, "tables of 2^32 rows or more are not supported on one GPU");
  t->n_rows = total;
  t->cols.resize(ncol);

  std::vector<uint8_t> hbits;
  for (int64_t ci = 0; ci < ncol; ++ci) {
    const uint8_t dt = t->fields[ci].dtype;
    // does any batch carry nulls?
    bool any_null = false;
    for (int64_t b = 0; b < n_batches && !any_null; ++b) {
      ArrowArray* a = batches[b].children[ci];
      if (dt == DT_NULL) break;
      if (a->n_buffers > 0 && a->buffers[0] != nullptr && a->null_count != 0) {
        // count over the rows the struct array exposes (its offset / length), not over the whole child: a sliced
        // batch may hide the child's nulls, and a longer child must not be read past the logical range
        const int64_t len = batches[b].length;
        any_null = len - count_bits((const uint8_t*)a->buffers[0], a->offset + batches[b].offset, len) > 0;
      }
    }
    t->cols[ci] = alloc_column(c, dt, total, false);  // owned by the table from here on: an error below frees it
    Column& col = t->cols[ci];
    const int w = dtype_width(dt);
    int64_t chars_total = 0;
    if (dt == DT_UTF8) {
      for (int64_t b = 0; b < n_batches; ++b) {
        ArrowArray* a = batches[b].children[ci];
        const int32_t* off = (const int32_t*)a->buffers[1];
        const int64_t o = a->offset + batches[b].offset;
        const int64_t len = batches[b].length;  // (the struct's length: a child may be longer than the struct that holds it)
        if (len > 0) chars_total += off[o + len] - off[o];
      }
      if (chars_total > 0x7ffffff0ll) throw_error(QGPU_ERR_UNSUPPORTED, "Utf8 column exceeds 2 GiB");
      col.chars = (uint8_t*)c->alloc(pad_bytes((size_t)chars_total));
      col.n_chars = chars_total;
    }
    std::vector<int32_t> hoff;
    if (dt == DT_UTF8) hoff.assign((size_t)total + 1, 0);
    if (dt == DT_BOOL) hbits.assign(bitmap_bytes(total), 0);
    int64_t row = 0, chars_at = 0;
    for (int64_t b = 0; b < n_batches; ++b) {
      ArrowArray* a = batches[b].children[ci];
      const int64_t o = a->offset + batches[b].offset;
      const int64_t len = batches[b].length;
      if (len == 0) continue;
      if (w > 0) {
        QG_CUDA(cudaMemcpyAsync((char*)col.data + row * w, (const char*)a->buffers[1] + o * w, (size_t)len * w,
                                cudaMemcpyHostToDevice, c->stream));
      } else if (dt == DT_BOOL) {
        copy_bits(hbits.data(), row, (const uint8_t*)a->buffers[1], o, len);
      } else if (dt == DT_UTF8) {
        const int32_t* off = (const int32_t*)a->buffers[1];
        const int32_t first = off[o];
        for (int64_t i = 0; i < len; ++i) hoff[row + i] = (int32_t)(chars_at + (off[o + i] - first));
        const int64_t nb = off[o + len] - first;
        if (nb > 0)
          QG_CUDA(cudaMemcpyAsync(col.chars + chars_at, (const char*)a->buffers[2] + first, (size_t)nb, cudaMemcpyHostToDevice,
                                  c->stream));
        chars_at += nb;
      }
      row += len;
    }
    if (dt == DT_UTF8) {
      hoff[total] = (int32_t)chars_at;
      QG_CUDA(cudaMemcpyAsync(col.data, hoff.data(), (size_t)(total + 1) * 4, cudaMemcpyHostToDevice, c->stream));
      QG_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (dt == DT_BOOL) {
      QG_CUDA(cudaMemcpyAsync(col.data, hbits.data(), (size_t)((total + 31) / 32) * 4, cudaMemcpyHostToDevice, c->stream));
      QG_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (any_null) {
      std::vector<uint8_t> hv(bitmap_bytes(total), 0);
      int64_t r = 0, nulls = 0;
      for (int64_t b = 0; b < n_batches; ++b) {
        ArrowArray* a = batches[b].children[ci];
        const int64_t o = a->offset + batches[b].offset;
        const int64_t len = batches[b].length;
        if (a->n_buffers > 0 && a->buffers[0] != nullptr && a->null_count != 0) {
          copy_bits(hv.data(), r, (const uint8_t*)a->buffers[0], o, len);
        } else {
          for (int64_t i = 0; i < len; ++i) hv[(r + i) >> 3] |= (uint8_t)(1u << ((r + i) & 7));
        }
        r += len;
      }
      nulls = total - count_bits(hv.data(), 0, total);
      col.valid = (uint32_t*)c->alloc(bitmap_bytes(total));
      QG_CUDA(cudaMemcpyAsync(col.valid, hv.data(), bitmap_bytes(total), cudaMemcpyHostToDevice, c->stream));
      QG_CUDA(cudaStreamSynchronize(c->stream));
      col.null_count = nulls;
    }
  }
  QG_CUDA(cudaStreamSynchronize(c->stream));
  return t;
}

// ---------------------------------------------------------------- export
namespace {

// Host memory of exported batches.  Large buffers are PINNED (cudaHostAlloc) so the device -> host copies run at PCIe
// speed instead of through the driver's pageable staging (measured: 1.9 GB of join output took 0.9 s pageable), and
// are recycled through a process-wide pool: the Arrow release callback returns them, the next export reuses them
// (pinning is slow, ~0.3 ms per MB).  The pool outlives contexts because exported batches may.
class HostPool {
 public:
  static HostPool& get() {
    static HostPool* p = new HostPool();  // leaked on purpose: release callbacks may run at exit
    return *p;
  }
  static constexpr size_t kPinnedMin = 1 << 20;
  void* acquire(size_t bytes, bool* pinned) {
    if (bytes < kPinnedMin) {
      *pinned = false;
      return std::malloc(bytes ? bytes : 8);
    }
    const size_t cls = round_up(bytes);
    {
      std::lock_guard<std::mutex> g(mu_);
      auto it = free_.find(cls);
      if (it != free_.end() && !it->second.empty()) {
        void* m = it->second.back();
        it->second.pop_back();
        cached_ -= cls;
        live_[m] = cls;
        *pinned = true;
        return m;
      }
    }
    void* m = nullptr;
    if (cudaHostAlloc(&m, cls, cudaHostAllocPortable) != cudaSuccess) {
      cudaGetLastError();
      *pinned = false;
      return std::malloc(bytes);
    }
    std::lock_guard<std::mutex> g(mu_);
    live_[m] = cls;
    *pinned = true;
    return m;
  }
  void release(void* m, bool pinned) {
    if (!m) return;
    if (!pinned) {
      std::free(m);
      return;
    }
    std::lock_guard<std::mutex> g(mu_);
    auto it = live_.find(m);
    if (it == live_.end()) return;
    const size_t cls = it->second;
    live_.erase(it);
    if (cached_ + cls > kMaxCached) {
      cudaFreeHost(m);
      return;
    }
    free_[cls].push_back(m);
    cached_ += cls;
  }

 private:
  static constexpr size_t kMaxCached = (size_t)16 << 30;
  static size_t round_up(size_t b) { return (b + ((size_t)2 << 20) - 1) / ((size_t)2 << 20) * ((size_t)2 << 20); }
  std::mutex mu_;
  std::map<size_t, std::vector<void*>> free_;
  std::map<void*, size_t> live_;
  size_t cached_ = 0;
};

struct ExportPrivate {
  std::vector<std::pair<void*, bool>> allocs;  // (buffer, pinned)
  std::vector<const void*> buffers;              // top-level struct buffers
  std::vector<ArrowArray*> children;
  std::vector<std::vector<const void*>> child_buffers;
  std::vector<ArrowArray> child_arrays;
};

void release_array(ArrowArray* a) {
  if (!a || !a->release) return;
  ExportPrivate* p = (ExportPrivate*)a->private_data;
  if (p) {
    for (auto& m : p->allocs) HostPool::get().release(m.first, m.second);
    delete p;
  }
  a->release = nullptr;
}
void release_child(ArrowArray* a) { a->release = nullptr; }

struct SchemaPrivate {
  std::vector<ArrowSchema> child_schemas;
  std::vector<ArrowSchema*> children;
  std::vector<std::string> names;
};
void release_schema(ArrowSchema* s) {
  if (!s || !s->release) return;
  delete (SchemaPrivate*)s->private_data;
  s->release = nullptr;
}
void release_child_schema(ArrowSchema* s) { s->release = nullptr; }

}  // namespace

void export_batch(Ctx* c, const Table& t, int64_t index, ArrowArray* out, ArrowSchema* out_schema) {
  const int64_t nb = (int64_t)t.batch_offsets.size() - 1;
  if (index < 0 || index >= nb) throw_error(QGPU_ERR_INVALID, "batch index %lld out of range (%lld)", (long long)index, (long long)nb);
  const int64_t r0 = t.batch_offsets[index], r1 = t.batch_offsets[index + 1], len = r1 - r0;
  const size_t ncol = t.cols.size();
  c->sync();

  auto* P = new ExportPrivate();
  P->child_arrays.resize(ncol);
  P->child_buffers.resize(ncol);
  P->children.resize(ncol);
  bool last_pinned = false;
  auto host_alloc = [&](size_t bytes) {
    void* m = HostPool::get().acquire(bytes, &last_pinned);
    if (!m) throw_error(QGPU_ERR_CUDA, "host allocation of %zu bytes failed", bytes);
    P->allocs.push_back({m, last_pinned});
    return m;
  };
  // device -> host: asynchronous into pinned buffers (all columns queue up, one sync at the end), blocking otherwise
  auto d2h = [&](void* dst, const void* src, size_t bytes, bool pinned) {
    if (pinned) QG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    else QG_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  };
  try {
    for (size_t ci = 0; ci < ncol; ++ci) {
      const Column& col = t.cols[ci];
      ArrowArray& a = P->child_arrays[ci];
      memset(&a, 0, sizeof a);
      a.length = len;
      a.offset = 0;
      a.release = release_child;
      std::vector<const void*>& bufs = P->child_buffers[ci];
      // validity
      const void* vbuf = nullptr;
      int64_t nulls = 0;
      auto fetch_bits = [&](const uint32_t* dbits) -> uint8_t* {
        const int64_t w0 = r0 >> 5, w1 = (r1 + 31) >> 5;
        std::vector<uint32_t> tmp((size_t)(w1 - w0) + 1, 0);
        if (w1 > w0) QG_CUDA(cudaMemcpy(tmp.data(), dbits + w0, (size_t)(w1 - w0) * 4, cudaMemcpyDeviceToHost));
        uint8_t* hb = (uint8_t*)host_alloc((size_t)((len + 63) / 64) * 8 + 8);
        memset(hb, 0, (size_t)((len + 63) / 64) * 8 + 8);
        copy_bits(hb, 0, (const uint8_t*)tmp.data(), r0 - (w0 << 5), len);
        return hb;
      };
      if (col.dtype == DT_NULL) {
        nulls = len;
      } else if (col.valid) {
        uint8_t* hb = fetch_bits(col.valid);
        nulls = len - count_bits(hb, 0, len);
        vbuf = hb;
      }
      a.null_count = nulls;
      const int w = dtype_width(col.dtype);
      if (col.dtype == DT_NULL) {
        a.n_buffers = 0;
      } else if (w > 0) {
        void* hd = host_alloc((size_t)len * w);
        if (len > 0) d2h(hd, (const char*)col.data + r0 * w, (size_t)len * w, last_pinned);
        bufs = {vbuf, hd};
        a.n_buffers = 2;
      } else if (col.dtype == DT_BOOL) {
        uint8_t* hb = fetch_bits((const uint32_t*)col.data);
        bufs = {vbuf, hb};
        a.n_buffers = 2;
      } else if (col.dtype == DT_UTF8) {
        int32_t* ho = (int32_t*)host_alloc((size_t)(len + 1) * 4);
        QG_CUDA(cudaMemcpy(ho, (const int32_t*)col.data + r0, (size_t)(len + 1) * 4, cudaMemcpyDeviceToHost));
        const int32_t first = ho[0];
        const int64_t nbytes = ho[len] - first;
        char* hc = (char*)host_alloc((size_t)nbytes);
        if (nbytes > 0) QG_CUDA(cudaMemcpy(hc, col.chars + first, (size_t)nbytes, cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i <= len; ++i) ho[i] -= first;
        bufs = {vbuf, ho, hc};
        a.n_buffers = 3;
      }
      a.buffers = bufs.empty() ? nullptr : bufs.data();
      P->children[ci] = &a;
    }
  } catch (...) {
    cudaStreamSynchronize(c->stream);
    for (auto& m : P->allocs) HostPool::get().release(m.first, m.second);
    delete P;
    throw;
  }
  QG_CUDA(cudaStreamSynchronize(c->stream));  // the asynchronous copies into pinned buffers
  memset(out, 0, sizeof *out);
  out->length = len;
  out->null_count = 0;
  out->offset = 0;
  P->buffers = {nullptr};
  out->n_buffers = 1;
  out->buffers = P->buffers.data();
  out->n_children = (int64_t)ncol;
  out->children = P->children.data();
  out->release = release_array;
  out->private_data = P;

  if (out_schema) {
    auto* S = new SchemaPrivate();
    S->child_schemas.resize(ncol);
    S->children.resize(ncol);
    S->names.resize(ncol);
    for (size_t ci = 0; ci < ncol; ++ci) {
      ArrowSchema& s = S->child_schemas[ci];
      memset(&s, 0, sizeof s);
      S->names[ci] = t.fields[ci].name;
      s.format = format_of_dtype(t.cols[ci].dtype);
      s.name = S->names[ci].c_str();
      s.flags = t.fields[ci].nullable ? ARROW_FLAG_NULLABLE : 0;
      s.release = release_child_schema;
      S->children[ci] = &s;
    }
    memset(out_schema, 0, sizeof *out_schema);
    out_schema->format = "+s";
    out_schema->name = "";
    out_schema->n_children = (int64_t)ncol;
    out_schema->children = S->children.data();
    out_schema->release = release_schema;
    out_schema->private_data = S;
  }
}

}  // namespace qgpu
