// SpectraST binary spectral libraries (".splib") straight into peak-store (CSR) arrays.
//
// Replaces the per-spectrum SplibParser.read_spectrum of reference parsers.pyx:41-186 (one Python
// MsmsSpectrum + one FragmentAnnotation object per peak) for library ingestion:
// two passes over the memory-mapped file — count, then fill caller-owned arrays — with the same field
// semantics as the reference parser:
//   header   : skip 8 bytes (two int32 versions), one line (file name), int32 k, k preamble lines (:94-99)
//   spectrum : uint32 identifier | line "n.PEPTIDE.c/charge ..." | float64 precursor m/z | line (status) |
//              uint32 num_peaks | num_peaks x (float64 m/z, float64 intensity, line annotation, line info) |
//              line comment; decoy <=> the comment contains " Remark=DECOY_" (:117-142)
//   peptide  : the text between the first and the second '.' of the name line (:122-124)
//   charge   : the integer after the first '/' behind the peptide (:125-128)
//   peaks    : m/z and intensity narrowed to float32 (:138-139)
//   peak charge (parse_annotation :160-186, what spectrum_match.pyx:74-79 later reads as
//              annotation[i].charge): ions a / b / y only; "<type><index>" directly followed by '/' -> 1;
//              followed by '^<z>' -> z; anything else (neutral losses, isotopes, other ions) -> no
//              annotation, stored as 0.
// Host code only (no GPU work): the arrays feed solo_process_spectra / solo_load_library.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstdlib>

#include "solo_common.cuh"

namespace solo {

namespace {

struct Splib {
    const char *p = nullptr;
    size_t size = 0, pos = 0;
    int fd = -1;
    const char *path = "";
    ~Splib() {
        if (p) munmap((void *)p, size);
        if (fd >= 0) close(fd);
    }
    void open(const char *path_) {
        path = path_;
        fd = ::open(path, O_RDONLY);
        SOLO_REQUIRE(fd >= 0, SOLO_EINVAL, "cannot open spectral library '%s': %s", path, strerror(errno));
        struct stat st;
        SOLO_REQUIRE(fstat(fd, &st) == 0 && st.st_size > 12, SOLO_EINVAL, "'%s' is not a .splib file", path);
        size = (size_t)st.st_size;
        void *m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        SOLO_REQUIRE(m != MAP_FAILED, SOLO_EINVAL, "cannot map '%s': %s", path, strerror(errno));
        p = (const char *)m;
    }
    void need(size_t n) {
        SOLO_REQUIRE(n <= size - pos, SOLO_EINVAL, "'%s' is truncated at byte %zu (need %zu more of %zu)", path, pos, n, size);
    }
    uint32_t u32() {
        need(4);
        uint32_t v;
        memcpy(&v, p + pos, 4);
        pos += 4;
        return v;
    }
    double f64() {
        need(8);
        double v;
        memcpy(&v, p + pos, 8);
        pos += 8;
        return v;
    }
    // [begin, end) of the line starting at pos, '\n' excluded; pos moves behind the '\n'
    void line(const char *&b, const char *&e) {
        b = p + pos;
        const void *nl = memchr(b, '\n', size - pos);
        e = nl ? (const char *)nl : p + size;
        pos = (size_t)(e - p) + (nl ? 1 : 0);
    }
    void skip_line() {
        const char *b, *e;
        line(b, e);
    }
    void seek_first_spectrum() {
        pos = 8;
        skip_line();
        uint32_t k = u32();
        for (uint32_t i = 0; i < k; ++i) skip_line();
    }
};

int parse_int(const char *b, const char *e) {
    int v = 0;
    bool any = false;
    while (b < e && *b >= '0' && *b <= '9') {
        v = v * 10 + (*b - '0');
        ++b;
        any = true;
    }
    return any ? v : -1;
}

// parsers.pyx:160-186
int annotation_charge(const char *b, const char *e) {
    if (b >= e || !(*b == 'a' || *b == 'b' || *b == 'y')) return 0;
    const char *q = b + 1;
    while (q < e && *q >= '0' && *q <= '9') ++q;   // find_first_not_of digits
    if (q == b + 1) return 0;                      // no ion index (stoi would throw in the reference)
    const char *slash = (const char *)memchr(q, '/', (size_t)(e - q));
    if (slash == q) return 1;
    if (q < e && *q == '^') {
        const int z = parse_int(q + 1, slash ? slash : e);
        return z > 0 && z < 256 ? z : 0;
    }
    return 0;
}

struct Sink {  // null members: counting pass
    uint32_t *id = nullptr;
    double *prec_mz = nullptr;
    int32_t *prec_charge = nullptr;
    uint8_t *is_decoy = nullptr;
    int64_t *file_offset = nullptr, *peak_off = nullptr, *pep_off = nullptr;
    float *mz = nullptr, *inten = nullptr;
    uint8_t *peak_charge = nullptr;
    char *pep = nullptr;
};

void walk(Splib &f, Sink *s, int64_t &n_spectra, int64_t &n_peaks, int64_t &n_pep) {
    f.seek_first_spectrum();
    n_spectra = n_peaks = n_pep = 0;
    static const char kDecoy[] = " Remark=DECOY_";
    while (f.pos < f.size) {
        const int64_t at = (int64_t)f.pos;
        const uint32_t ident = f.u32();
        const char *b, *e;
        f.line(b, e);
        const char *d1 = (const char *)memchr(b, '.', (size_t)(e - b));
        SOLO_REQUIRE(d1, SOLO_EINVAL, "'%s': spectrum at byte %lld has no 'n.PEPTIDE.c/z' name", f.path, (long long)at);
        const char *pb = d1 + 1;
        const char *d2 = (const char *)memchr(pb, '.', (size_t)(e - pb));
        SOLO_REQUIRE(d2, SOLO_EINVAL, "'%s': spectrum at byte %lld: malformed name", f.path, (long long)at);
        const char *sl = (const char *)memchr(d2, '/', (size_t)(e - d2));
        const int z = sl ? parse_int(sl + 1, e) : -1;
        SOLO_REQUIRE(z >= 0, SOLO_EINVAL, "'%s': spectrum at byte %lld: no precursor charge", f.path, (long long)at);
        const double pm = f.f64();
        f.skip_line();  // status
        const uint32_t np = f.u32();
        SOLO_REQUIRE((size_t)np * 18 <= f.size - f.pos, SOLO_EINVAL, "'%s': spectrum at byte %lld claims %u peaks",
                     f.path, (long long)at, np);
        if (s) {
            s->id[n_spectra] = ident;
            s->prec_mz[n_spectra] = pm;
            s->prec_charge[n_spectra] = z;
            s->file_offset[n_spectra] = at;
            s->peak_off[n_spectra] = n_peaks;
            s->pep_off[n_spectra] = n_pep;
            memcpy(s->pep + n_pep, pb, (size_t)(d2 - pb));
        }
        for (uint32_t i = 0; i < np; ++i) {
            const double m = f.f64(), in = f.f64();
            const char *ab, *ae;
            f.line(ab, ae);
            f.skip_line();
            if (s) {
                s->mz[n_peaks + i] = (float)m;
                s->inten[n_peaks + i] = (float)in;
                s->peak_charge[n_peaks + i] = (uint8_t)annotation_charge(ab, ae);
            }
        }
        f.line(b, e);  // comment
        if (s) {
            bool decoy = false;
            for (const char *q = b; q + sizeof(kDecoy) - 1 <= e; ++q)
                if (memcmp(q, kDecoy, sizeof(kDecoy) - 1) == 0) {
                    decoy = true;
                    break;
                }
            s->is_decoy[n_spectra] = decoy;
        }
        n_peaks += np;
        n_pep += d2 - pb;
        ++n_spectra;
    }
    if (s) {
        s->peak_off[n_spectra] = n_peaks;
        s->pep_off[n_spectra] = n_pep;
    }
}

}  // namespace

void splib_count(const char *path, int64_t *n_spectra, int64_t *n_peaks, int64_t *n_peptide_bytes) {
    Splib f;
    f.open(path);
    walk(f, nullptr, *n_spectra, *n_peaks, *n_peptide_bytes);
}

void splib_read(const char *path, int64_t n_spectra, int64_t n_peaks, int64_t n_peptide_bytes, uint32_t *id,
                double *prec_mz, int32_t *prec_charge, uint8_t *is_decoy, int64_t *file_offset, int64_t *peak_off,
                float *mz, float *inten, uint8_t *peak_charge, int64_t *pep_off, char *pep) {
    Splib f;
    f.open(path);
    int64_t a, b, c;
    walk(f, nullptr, a, b, c);
    SOLO_REQUIRE(a == n_spectra && b == n_peaks && c == n_peptide_bytes, SOLO_EINVAL,
                 "'%s' holds %lld spectra / %lld peaks / %lld peptide bytes, the buffers were sized for %lld / %lld / %lld",
                 path, (long long)a, (long long)b, (long long)c, (long long)n_spectra, (long long)n_peaks,
                 (long long)n_peptide_bytes);
    Sink s{id, prec_mz, prec_charge, is_decoy, file_offset, peak_off, pep_off, mz, inten, peak_charge, pep};
    walk(f, &s, a, b, c);
}

}  // namespace solo
