// Faiss index files (".idxann") in and out of the device IVF index.
//
// Replaces faiss.write_index / faiss.read_index at reference spectral_library.py:181 and :490 for
// the one index type the reference builds (IndexIVFFlat over an IndexFlatIP quantizer, :167-176).
// Faiss itself is absent here; the byte layout below restates the published index_write.cpp /
// index_read.cpp serialisation of that type (little-endian, fields in this order):
//
//   "IwFl"                                  fourcc of IndexIVFFlat
//   d i32 | ntotal i64 | 1<<20 i64 | 1<<20 i64 | is_trained u8 | metric i32 [| metric_arg f32 if metric > 1]
//   nlist u64 | nprobe u64
//   quantizer: "IxFI" (IP; "IxF2" = L2, "IxFl" = legacy) + the same header + count u64 + count f32
//   direct map: type u8 | n u64 | n * i64 [| n u64 | n * (i64, i64) if type == 2 (hash table)]
//   "ilar" | nlist u64 | code_size u64
//   "full" | nlist u64 | nlist * u64 sizes      or      "sprs" | 2m u64 | m * (list u64, size u64)
//   for every non-empty list, in list order: size * code_size bytes of codes, then size * i64 ids
//
// The reader maps the file, keeps the inverted lists exactly as stored (no re-assignment) and hands
// the rows to the device store in id order; the writer streams the lists back from the device.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cmath>

#include "ivf.cuh"

namespace solo {

namespace {

struct MappedFile {
    const uint8_t *p = nullptr;
    size_t size = 0;
    int fd = -1;
    ~MappedFile() {
        if (p) munmap((void *)p, size);
        if (fd >= 0) close(fd);
    }
    void open(const char *path) {
        fd = ::open(path, O_RDONLY);
        SOLO_REQUIRE(fd >= 0, SOLO_EINVAL, "cannot open index file '%s': %s", path, strerror(errno));
        struct stat st;
        SOLO_REQUIRE(fstat(fd, &st) == 0, SOLO_EINVAL, "cannot stat '%s': %s", path, strerror(errno));
        size = (size_t)st.st_size;
        SOLO_REQUIRE(size >= 4, SOLO_EINVAL, "'%s' is not a Faiss index (only %zu bytes)", path, size);
        void *m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        SOLO_REQUIRE(m != MAP_FAILED, SOLO_EINVAL, "cannot map '%s': %s", path, strerror(errno));
        p = (const uint8_t *)m;
    }
};

struct Cursor {
    const uint8_t *p;
    size_t size, pos = 0;
    const char *path;
    void need(size_t n, const char *what) {
        SOLO_REQUIRE(n <= size - pos, SOLO_EINVAL, "'%s' is truncated: %s needs %zu bytes at offset %zu of %zu", path,
                     what, n, pos, size);
    }
    template <typename T>
    T get(const char *what) {
        need(sizeof(T), what);
        T v;
        memcpy(&v, p + pos, sizeof(T));
        pos += sizeof(T);
        return v;
    }
    const uint8_t *skip(size_t n, const char *what) {
        need(n, what);
        const uint8_t *q = p + pos;
        pos += n;
        return q;
    }
    std::string fourcc(const char *what) {
        need(4, what);
        std::string s((const char *)p + pos, 4);
        pos += 4;
        return s;
    }
};

struct IndexHeader {
    int32_t d = 0;
    int64_t ntotal = 0;
    int is_trained = 0;
    int32_t metric = 0;
};

IndexHeader read_index_header(Cursor &c) {
    IndexHeader hd;
    hd.d = c.get<int32_t>("d");
    hd.ntotal = c.get<int64_t>("ntotal");
    c.get<int64_t>("reserved");
    c.get<int64_t>("reserved");
    hd.is_trained = c.get<uint8_t>("is_trained");
    hd.metric = c.get<int32_t>("metric_type");
    if (hd.metric > 1) c.get<float>("metric_arg");
    return hd;
}

// Everything read_index needs, pointing into the mapping.
struct ParsedIvfFlat {
    IndexHeader ivf, quant;
    std::string fourcc, quant_fourcc;
    uint64_t nlist = 0, nprobe = 0, code_size = 0;
    const uint8_t *centroids = nullptr;  // nlist * d float32 (unaligned)
    std::vector<uint64_t> list_size;
    std::vector<const uint8_t *> list_codes, list_ids;
    uint64_t nstored = 0, max_list_len = 0;
};

void parse_ivf_flat(Cursor &c, ParsedIvfFlat &out) {
    out.fourcc = c.fourcc("index fourcc");
    SOLO_REQUIRE(out.fourcc == "IwFl", SOLO_EINVAL,
                 "'%s': index type '%s' is not supported (the reference writes IndexIVFFlat, 'IwFl')", c.path,
                 out.fourcc.c_str());
    out.ivf = read_index_header(c);
    out.nlist = c.get<uint64_t>("nlist");
    out.nprobe = c.get<uint64_t>("nprobe");
    SOLO_REQUIRE(out.ivf.d > 0 && out.ivf.d <= 1536, SOLO_ECAPACITY, "'%s': d = %d outside [1, 1536]", c.path, out.ivf.d);
    SOLO_REQUIRE(out.nlist > 0 && out.nlist <= (uint64_t)IVF_MAX_NLIST, SOLO_ECAPACITY,
                 "'%s': nlist = %llu outside [1, %d]", c.path, (unsigned long long)out.nlist, IVF_MAX_NLIST);
    SOLO_REQUIRE(out.ivf.ntotal >= 0 && out.ivf.ntotal < (int64_t)0x7fffffff, SOLO_ECAPACITY, "'%s': ntotal = %lld",
                 c.path, (long long)out.ivf.ntotal);
    // coarse quantizer: a flat index holding the nlist centroids
    out.quant_fourcc = c.fourcc("quantizer fourcc");
    SOLO_REQUIRE(out.quant_fourcc == "IxFI" || out.quant_fourcc == "IxF2" || out.quant_fourcc == "IxFl", SOLO_EINVAL,
                 "'%s': quantizer type '%s' is not a flat index", c.path, out.quant_fourcc.c_str());
    out.quant = read_index_header(c);
    uint64_t nfloat = c.get<uint64_t>("quantizer vector count");
    SOLO_REQUIRE(out.quant.d == out.ivf.d && (uint64_t)out.quant.ntotal == out.nlist &&
                     nfloat == out.nlist * (uint64_t)out.ivf.d,
                 SOLO_EINVAL, "'%s': quantizer holds %lld x %d values (%llu), expected %llu x %d", c.path,
                 (long long)out.quant.ntotal, out.quant.d, (unsigned long long)nfloat, (unsigned long long)out.nlist,
                 out.ivf.d);
    out.centroids = c.skip(nfloat * sizeof(float), "centroids");
    // direct map (the reference never maintains one; skipped whatever its type)
    uint8_t dm_type = c.get<uint8_t>("direct map type");
    uint64_t dm_n = c.get<uint64_t>("direct map size");
    c.skip(dm_n * sizeof(int64_t), "direct map");
    if (dm_type == 2) {
        uint64_t hn = c.get<uint64_t>("direct map hash size");
        c.skip(hn * 2 * sizeof(int64_t), "direct map hash table");
    }
    // inverted lists
    std::string il = c.fourcc("inverted lists fourcc");
    SOLO_REQUIRE(il == "ilar", SOLO_EINVAL, "'%s': inverted lists of type '%s' are not supported (expected 'ilar')",
                 c.path, il.c_str());
    uint64_t il_nlist = c.get<uint64_t>("ilar nlist");
    out.code_size = c.get<uint64_t>("code_size");
    SOLO_REQUIRE(il_nlist == out.nlist, SOLO_EINVAL, "'%s': %llu inverted lists for nlist = %llu", c.path,
                 (unsigned long long)il_nlist, (unsigned long long)out.nlist);
    SOLO_REQUIRE(out.code_size == (uint64_t)out.ivf.d * sizeof(float), SOLO_EINVAL,
                 "'%s': code_size %llu is not d * 4 (IVF-Flat stores float32 rows)", c.path,
                 (unsigned long long)out.code_size);
    out.list_size.assign(out.nlist, 0);
    std::string lt = c.fourcc("list size encoding");
    uint64_t nv = c.get<uint64_t>("list size vector");
    if (lt == "full") {
        SOLO_REQUIRE(nv == out.nlist, SOLO_EINVAL, "'%s': %llu list sizes for nlist = %llu", c.path,
                     (unsigned long long)nv, (unsigned long long)out.nlist);
        for (uint64_t l = 0; l < out.nlist; ++l) out.list_size[l] = c.get<uint64_t>("list size");
    } else if (lt == "sprs") {
        SOLO_REQUIRE(nv % 2 == 0, SOLO_EINVAL, "'%s': odd sparse list-size vector", c.path);
        for (uint64_t i = 0; i < nv; i += 2) {
            uint64_t l = c.get<uint64_t>("list id"), n = c.get<uint64_t>("list size");
            SOLO_REQUIRE(l < out.nlist, SOLO_EINVAL, "'%s': list id %llu out of range", c.path, (unsigned long long)l);
            out.list_size[l] = n;
        }
    } else {
        SOLO_REQUIRE(false, SOLO_EINVAL, "'%s': unknown list size encoding '%s'", c.path, lt.c_str());
    }
    out.list_codes.assign(out.nlist, nullptr);
    out.list_ids.assign(out.nlist, nullptr);
    for (uint64_t l = 0; l < out.nlist; ++l) {
        uint64_t n = out.list_size[l];
        if (!n) continue;
        SOLO_REQUIRE(n <= c.size / out.code_size, SOLO_EINVAL, "'%s' is truncated: list %llu claims %llu rows", c.path,
                     (unsigned long long)l, (unsigned long long)n);
        out.list_codes[l] = c.skip(n * out.code_size, "list codes");
        out.list_ids[l] = c.skip(n * sizeof(int64_t), "list ids");
        out.nstored += n;
        out.max_list_len = std::max(out.max_list_len, n);
    }
}

// sparse rows [first, first + m) of the insertion-ordered store (or of the list-ordered store via
// `rows`) back to dense float32: the inverse of dense_fill_kernel, one warp per row
__global__ void __launch_bounds__(256)
rows_to_dense_kernel(const int32_t *__restrict__ rows, int64_t first, int64_t m, const int64_t *__restrict__ r_off,
                     const uint16_t *__restrict__ r_idx, const float *__restrict__ r_val, int d,
                     float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= m) return;
    const int64_t r = rows ? (int64_t)rows[first + i] : first + i;
    float *o = out + i * d;
    for (int j = lane; j < d; j += 32) o[j] = 0.f;
    __syncwarp();
    for (int64_t p = r_off[r] + lane; p < r_off[r + 1]; p += 32) o[r_idx[p]] = r_val[p];
}

// rows just appended with a placeholder list: take the list from the file; rows the store skipped
// (NaN) stay skipped
__global__ void apply_assignment_kernel(int32_t *__restrict__ row_list, const int32_t *__restrict__ given, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && row_list[i] >= 0) row_list[i] = given[i];
}

struct PinnedBuf {
    void *p = nullptr;
    ~PinnedBuf() {
        if (p) cudaFreeHost(p);
    }
    void alloc(size_t bytes) { SOLO_CUDA(cudaMallocHost(&p, std::max<size_t>(bytes, 16))); }
};

void put(FILE *f, const void *p, size_t n, const char *path) {
    SOLO_REQUIRE(n == 0 || fwrite(p, 1, n, f) == n, SOLO_EINVAL, "short write to '%s': %s", path, strerror(errno));
}
template <typename T>
void put1(FILE *f, T v, const char *path) {
    put(f, &v, sizeof(T), path);
}
void put_index_header(FILE *f, int32_t d, int64_t ntotal, const char *path) {
    put1<int32_t>(f, d, path);
    put1<int64_t>(f, ntotal, path);
    put1<int64_t>(f, (int64_t)1 << 20, path);
    put1<int64_t>(f, (int64_t)1 << 20, path);
    put1<uint8_t>(f, 1, path);   // is_trained
    put1<int32_t>(f, 0, path);   // METRIC_INNER_PRODUCT
}

}  // namespace

// Append `n` host rows with the lists given by the caller (list_of_row[i] in [0, nlist), or -1 to
// reserve the id without storing the row), instead of the arg-max-centroid assignment of add().
void ivf_add_assigned(solo_handle *h, IvfIndex &ix, const float *h_x, int64_t n, const int32_t *list_of_row) {
    SOLO_REQUIRE(ix.nlist > 0, SOLO_ESTATE, "index has no centroids (train or set_centroids first)");
    for (int64_t i = 0; i < n; ++i)
        SOLO_REQUIRE(list_of_row[i] >= -1 && list_of_row[i] < ix.nlist, SOLO_EINVAL,
                     "row %lld: list %d outside [-1, %d)", (long long)i, list_of_row[i], ix.nlist);
    DevBuf &xd = h->scratch[19], &given = h->scratch[20];
    const int64_t chunk = 1 << 17;
    xd.ensure((size_t)std::min<int64_t>(std::max<int64_t>(n, 1), chunk) * ix.dim * sizeof(float));
    given.ensure((size_t)std::min<int64_t>(std::max<int64_t>(n, 1), chunk) * sizeof(int32_t));
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        const int64_t m = std::min(chunk, n - r0), row0 = ix.ntotal;
        SOLO_CUDA(cudaMemcpyAsync(xd.p, h_x + r0 * ix.dim, (size_t)m * ix.dim * sizeof(float), cudaMemcpyHostToDevice,
                                  h->stream));
        SOLO_CUDA(cudaMemcpyAsync(given.p, list_of_row + r0, m * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        ivf_add_device(h, ix, xd.as<float>(), m, /*assign=*/false);
        apply_assignment_kernel<<<div_up(m, 256), 256, 0, h->stream>>>(ix.row_list.as<int32_t>() + row0,
                                                                      given.as<int32_t>(), m);
        SOLO_CUDA(cudaGetLastError());
        h->launches++;
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
    }
    ix.dirty = true;
}

void ivf_read_index(solo_handle *h, IvfIndex &ix, const char *path, int64_t *nprobe_out) {
    MappedFile mf;
    mf.open(path);
    Cursor c{mf.p, mf.size, 0, path};
    ParsedIvfFlat pf;
    parse_ivf_flat(c, pf);
    SOLO_REQUIRE(pf.ivf.metric == 0, SOLO_EINVAL,
                 "'%s': metric %d; only METRIC_INNER_PRODUCT (0) is implemented (spectral_library.py:175)", path,
                 pf.ivf.metric);
    const int d = pf.ivf.d;
    const int64_t ntotal = pf.ivf.ntotal;
    // ids -> (list, address of the row). The reference adds rows with sequential ids
    // (spectral_library.py:179), so the ids of a valid file are distinct and below ntotal.
    std::vector<int32_t> row_list((size_t)std::max<int64_t>(ntotal, 1), -1);
    std::vector<const uint8_t *> row_ptr((size_t)std::max<int64_t>(ntotal, 1), nullptr);
    for (uint64_t l = 0; l < pf.nlist; ++l) {
        for (uint64_t j = 0; j < pf.list_size[l]; ++j) {
            int64_t id;
            memcpy(&id, pf.list_ids[l] + j * sizeof(int64_t), sizeof id);
            SOLO_REQUIRE(id >= 0 && id < ntotal, SOLO_EINVAL,
                         "'%s': list %llu holds id %lld outside [0, ntotal = %lld): only sequential ids are supported",
                         path, (unsigned long long)l, (long long)id, (long long)ntotal);
            SOLO_REQUIRE(row_ptr[id] == nullptr, SOLO_EINVAL, "'%s': id %lld is stored twice", path, (long long)id);
            row_ptr[id] = pf.list_codes[l] + j * pf.code_size;
            row_list[id] = (int32_t)l;
        }
    }
    {   // centroids may sit at an odd offset in the mapping
        std::vector<float> cent((size_t)pf.nlist * d);
        memcpy(cent.data(), pf.centroids, cent.size() * sizeof(float));
        ivf_set_centroids(h, ix, cent.data(), (int)pf.nlist, d);
    }
    const int64_t chunk = 1 << 16;
    PinnedBuf stage;
    stage.alloc((size_t)std::min<int64_t>(std::max<int64_t>(ntotal, 1), chunk) * d * sizeof(float));
    float *sx = (float *)stage.p;
    for (int64_t r0 = 0; r0 < ntotal; r0 += chunk) {
        const int64_t m = std::min(chunk, ntotal - r0);
        for (int64_t i = 0; i < m; ++i) {
            float *dst = sx + i * d;
            if (row_ptr[r0 + i]) memcpy(dst, row_ptr[r0 + i], (size_t)d * sizeof(float));
            else memset(dst, 0, (size_t)d * sizeof(float));  // id not stored in any list: reserved, list -1
        }
        ivf_add_assigned(h, ix, sx, m, row_list.data() + r0);
    }
    if (nprobe_out) *nprobe_out = (int64_t)pf.nprobe;
}

void ivf_write_index(solo_handle *h, IvfIndex &ix, const char *path, int64_t nprobe) {
    SOLO_REQUIRE(ix.nlist > 0, SOLO_ESTATE, "index has no centroids (nothing to write)");
    SOLO_REQUIRE(ix.owned.empty(), SOLO_ESTATE, "a list-sharded index holds only its own lists; write it before sharding");
    ivf_finalize(h, ix);
    const int d = ix.dim, nlist = ix.nlist;
    FILE *f = fopen(path, "wb");
    SOLO_REQUIRE(f != nullptr, SOLO_EINVAL, "cannot create '%s': %s", path, strerror(errno));
    struct Closer {
        FILE *f;
        ~Closer() {
            if (f) fclose(f);
        }
    } closer{f};
    std::vector<char> iobuf(8 << 20);
    setvbuf(f, iobuf.data(), _IOFBF, iobuf.size());
    put(f, "IwFl", 4, path);
    put_index_header(f, d, ix.ntotal, path);
    put1<uint64_t>(f, (uint64_t)nlist, path);
    put1<uint64_t>(f, (uint64_t)std::max<int64_t>(nprobe, 1), path);
    {   // quantizer
        put(f, "IxFI", 4, path);
        put_index_header(f, d, nlist, path);
        std::vector<float> cent((size_t)nlist * d);
        SOLO_CUDA(cudaMemcpyAsync(cent.data(), ix.cent.p, cent.size() * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
        put1<uint64_t>(f, (uint64_t)cent.size(), path);
        put(f, cent.data(), cent.size() * sizeof(float), path);
    }
    put1<uint8_t>(f, 0, path);   // direct map: none
    put1<uint64_t>(f, 0, path);
    put(f, "ilar", 4, path);
    put1<uint64_t>(f, (uint64_t)nlist, path);
    put1<uint64_t>(f, (uint64_t)d * sizeof(float), path);
    const std::vector<int64_t> &off = ix.h_list_off;
    int64_t non0 = 0;
    for (int l = 0; l < nlist; ++l) non0 += off[l + 1] > off[l];
    if (non0 > nlist / 2) {
        put(f, "full", 4, path);
        put1<uint64_t>(f, (uint64_t)nlist, path);
        for (int l = 0; l < nlist; ++l) put1<uint64_t>(f, (uint64_t)(off[l + 1] - off[l]), path);
    } else {
        put(f, "sprs", 4, path);
        put1<uint64_t>(f, (uint64_t)(2 * non0), path);
        for (int l = 0; l < nlist; ++l)
            if (off[l + 1] > off[l]) {
                put1<uint64_t>(f, (uint64_t)l, path);
                put1<uint64_t>(f, (uint64_t)(off[l + 1] - off[l]), path);
            }
    }
    const int64_t ns = ix.nstored;
    std::vector<int32_t> ids((size_t)std::max<int64_t>(ns, 1));
    if (ns) SOLO_CUDA(cudaMemcpyAsync(ids.data(), ix.list_ids.p, ns * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    SOLO_CUDA(cudaStreamSynchronize(h->stream));
    // whole lists at a time, as many as fit the staging buffer
    const int64_t cap = std::max<int64_t>(ix.max_list_len, 1 << 15);
    DevBuf &dense = h->scratch[19];
    dense.ensure((size_t)cap * d * sizeof(float));
    PinnedBuf stage;
    stage.alloc((size_t)cap * d * sizeof(float));
    std::vector<int64_t> id64;
    for (int la = 0; la < nlist;) {
        int lb = la + 1;
        while (lb < nlist && off[lb + 1] - off[la] <= cap) ++lb;
        const int64_t p0 = off[la], m = off[lb] - off[la];
        if (m) {
            rows_to_dense_kernel<<<div_up(m, 8), 256, 0, h->stream>>>(ix.list_ids.as<int32_t>(), p0, m,
                                                                     ix.row_off.as<int64_t>(), ix.row_idx.as<uint16_t>(),
                                                                     ix.row_val.as<float>(), d, dense.as<float>());
            SOLO_CUDA(cudaGetLastError());
            h->launches++;
            SOLO_CUDA(cudaMemcpyAsync(stage.p, dense.p, (size_t)m * d * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
            SOLO_CUDA(cudaStreamSynchronize(h->stream));
            for (int l = la; l < lb; ++l) {
                const int64_t n = off[l + 1] - off[l];
                if (!n) continue;
                put(f, (const float *)stage.p + (off[l] - p0) * d, (size_t)n * d * sizeof(float), path);
                id64.resize(n);
                for (int64_t j = 0; j < n; ++j) id64[j] = ids[off[l] + j];
                put(f, id64.data(), (size_t)n * sizeof(int64_t), path);
            }
        }
        la = lb;
    }
    SOLO_REQUIRE(fflush(f) == 0, SOLO_EINVAL, "cannot flush '%s': %s", path, strerror(errno));
    closer.f = nullptr;
    SOLO_REQUIRE(fclose(f) == 0, SOLO_EINVAL, "cannot close '%s': %s", path, strerror(errno));
}

// dense float32 copy of stored rows [row0, row0 + n) in insertion order (rows that were skipped come back as zeros)
void ivf_reconstruct(solo_handle *h, IvfIndex &ix, int64_t row0, int64_t n, float *h_out) {
    SOLO_REQUIRE(row0 >= 0 && n >= 0 && row0 + n <= ix.ntotal, SOLO_EINVAL, "rows [%lld, %lld) outside [0, %lld)",
                 (long long)row0, (long long)(row0 + n), (long long)ix.ntotal);
    DevBuf &dense = h->scratch[19];
    const int64_t chunk = 1 << 15;
    dense.ensure((size_t)std::min<int64_t>(std::max<int64_t>(n, 1), chunk) * ix.dim * sizeof(float));
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        const int64_t m = std::min(chunk, n - r0);
        rows_to_dense_kernel<<<div_up(m, 8), 256, 0, h->stream>>>(nullptr, row0 + r0, m, ix.row_off.as<int64_t>(),
                                                                 ix.row_idx.as<uint16_t>(), ix.row_val.as<float>(),
                                                                 ix.dim, dense.as<float>());
        SOLO_CUDA(cudaGetLastError());
        h->launches++;
        SOLO_CUDA(cudaMemcpyAsync(h_out + r0 * ix.dim, dense.p, (size_t)m * ix.dim * sizeof(float),
                                  cudaMemcpyDeviceToHost, h->stream));
        SOLO_CUDA(cudaStreamSynchronize(h->stream));
    }
}

void idxann_inspect(const char *path, solo_idxann_info *info) {
    MappedFile mf;
    mf.open(path);
    Cursor c{mf.p, mf.size, 0, path};
    ParsedIvfFlat pf;
    parse_ivf_flat(c, pf);
    memset(info, 0, sizeof *info);
    info->d = pf.ivf.d;
    info->metric = pf.ivf.metric;
    info->is_trained = pf.ivf.is_trained;
    info->ntotal = pf.ivf.ntotal;
    info->nlist = (int64_t)pf.nlist;
    info->nprobe = (int64_t)pf.nprobe;
    info->code_size = (int64_t)pf.code_size;
    info->nstored = (int64_t)pf.nstored;
    info->max_list_len = (int64_t)pf.max_list_len;
    info->bytes_parsed = (int64_t)c.pos;
    memcpy(info->fourcc, pf.fourcc.data(), 4);
    memcpy(info->quantizer_fourcc, pf.quant_fourcc.data(), 4);
}

}  // namespace solo
