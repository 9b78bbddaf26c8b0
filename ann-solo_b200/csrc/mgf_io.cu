// Mascot Generic Format query files straight into peak-store (CSR) arrays.
//
// Replaces reader.read_mgf (reference reader.py:868-911, one pyteomics dict + one MsmsSpectrum per
// spectrum) for the query side of SURVEY.md §8f N3: two passes over the memory-mapped text — count,
// then fill caller-owned arrays that feed solo_process_spectra (K0) and the staged search.
// pyteomics is absent from the reference tree; the format semantics restated here are the ones the
// reference relies on (parity unpinned beyond the reference's own reader and its test's MassIVE-KB
// style entries, src/tests/query_reader_test.py:41-66):
//   * a spectrum is the text between a "BEGIN IONS" and the next "END IONS" line; everything outside is
//     ignored (file-level parameters); blank lines and lines starting with # ; ! / are comments
//   * inside, a line with '=' is a parameter, key case-insensitive: TITLE, SCAN(S), PEPMASS (first
//     number = precursor m/z), CHARGE ("2+", "3-", "2+ and 3+": the first entry, signed), RTINSECONDS,
//     SEQ, DECOY (presence); any other line is a peak "m/z intensity [charge]"
//   * identifier = TITLE, else SCAN/SCANS, else the 1-based index of the spectrum (the reference raises
//     KeyError there); precursor charge 0 = not given (None in the reference: the search then tries
//     2+ and 3+, spectral_library.py:219-223); retention time NaN = not given
//   * peaks are returned m/z-ascending (spectrum_utils.MsmsSpectrum orders them on construction), m/z
//     as float64, intensity narrowed to float32
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <charconv>
#include <cmath>
#include <numeric>

#include "solo_common.cuh"

namespace solo {

namespace {

struct TextFile {
    const char *p = nullptr;
    size_t size = 0;
    int fd = -1;
    ~TextFile() {
        if (p && size) munmap((void *)p, size);
        if (fd >= 0) close(fd);
    }
    void open(const char *path) {
        fd = ::open(path, O_RDONLY);
        SOLO_REQUIRE(fd >= 0, SOLO_EINVAL, "cannot open peak file '%s': %s", path, strerror(errno));
        struct stat st;
        SOLO_REQUIRE(fstat(fd, &st) == 0, SOLO_EINVAL, "cannot stat '%s': %s", path, strerror(errno));
        size = (size_t)st.st_size;
        if (size == 0) return;
        void *m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        SOLO_REQUIRE(m != MAP_FAILED, SOLO_EINVAL, "cannot map '%s': %s", path, strerror(errno));
        p = (const char *)m;
    }
};

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\f' || c == '\v'; }

void strip(const char *&b, const char *&e) {
    while (b < e && is_space(*b)) ++b;
    while (e > b && is_space(e[-1])) --e;
}

bool ieq(const char *b, const char *e, const char *word) {
    size_t n = strlen(word);
    if ((size_t)(e - b) != n) return false;
    for (size_t i = 0; i < n; ++i) {
        char c = b[i];
        if (c >= 'a' && c <= 'z') c = (char)(c - 32);
        if (c != word[i]) return false;
    }
    return true;
}

// first whitespace-separated token of [b, e) as a double; false when it is not a number
bool first_number(const char *b, const char *e, double &v, const char **rest = nullptr) {
    while (b < e && is_space(*b)) ++b;
    const char *t = b;
    while (t < e && !is_space(*t)) ++t;
    if (b < t && *b == '+') ++b;  // from_chars rejects a leading plus, float() accepts it
    auto r = std::from_chars(b, t, v);
    if (r.ec != std::errc() || r.ptr != t) return false;
    if (rest) *rest = t;
    return true;
}

// "2+", "3-", "2", "2+ and 3+" -> signed first entry; 0 when unparsable
int parse_charge(const char *b, const char *e) {
    while (b < e && is_space(*b)) ++b;
    int sign = 1;
    if (b < e && (*b == '+' || *b == '-')) {
        sign = *b == '-' ? -1 : 1;
        ++b;
    }
    int v = 0;
    bool any = false;
    while (b < e && *b >= '0' && *b <= '9') {
        v = v * 10 + (*b - '0');
        ++b;
        any = true;
    }
    if (!any) return 0;
    if (b < e && *b == '-') sign = -1;
    return sign * v;
}

struct MgfSink {  // null: counting pass
    double *prec_mz;
    int32_t *prec_charge;
    double *rt;
    uint8_t *is_decoy;
    int64_t *peak_off;
    double *mz;
    float *inten;
    int64_t *title_off;
    char *titles;
    int64_t *seq_off;
    char *seqs;
};

void walk_mgf(const TextFile &f, const char *path, MgfSink *s, int64_t &n_spec, int64_t &n_peaks, int64_t &n_title,
              int64_t &n_seq) {
    n_spec = n_peaks = n_title = n_seq = 0;
    const char *p = f.p, *end = f.p + f.size;
    bool inside = false;
    int64_t first_peak = 0;
    const char *title_b = nullptr, *title_e = nullptr, *scan_b = nullptr, *scan_e = nullptr, *seq_b = nullptr,
               *seq_e = nullptr;
    double pm = 0.0, rt = NAN;
    int z = 0;
    bool decoy = false, have_pm = false;
    std::vector<int32_t> order;
    std::vector<double> tmz;
    std::vector<float> tin;
    int64_t line_no = 0;
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *b = p, *e = nl ? nl : end;
        p = nl ? nl + 1 : end;
        ++line_no;
        strip(b, e);
        if (b == e) continue;
        if (!inside) {
            if (ieq(b, e, "BEGIN IONS")) {
                inside = true;
                first_peak = n_peaks;
                title_b = title_e = scan_b = scan_e = seq_b = seq_e = nullptr;
                pm = 0.0;
                rt = NAN;
                z = 0;
                decoy = have_pm = false;
            }
            continue;
        }
        if (ieq(b, e, "END IONS")) {
            SOLO_REQUIRE(have_pm, SOLO_EINVAL, "'%s': spectrum ending at line %lld has no PEPMASS", path, (long long)line_no);
            char idx_buf[24];
            const char *ib = title_b ? title_b : scan_b, *ie = title_b ? title_e : scan_e;
            if (!ib) {  // neither TITLE nor SCAN: 1-based index
                int len = snprintf(idx_buf, sizeof idx_buf, "%lld", (long long)(n_spec + 1));
                ib = idx_buf;
                ie = idx_buf + len;
            }
            if (s) {
                s->prec_mz[n_spec] = pm;
                s->prec_charge[n_spec] = z;
                s->rt[n_spec] = rt;
                s->is_decoy[n_spec] = decoy;
                s->peak_off[n_spec] = first_peak;
                s->title_off[n_spec] = n_title;
                s->seq_off[n_spec] = n_seq;
                memcpy(s->titles + n_title, ib, (size_t)(ie - ib));
                if (seq_b) memcpy(s->seqs + n_seq, seq_b, (size_t)(seq_e - seq_b));
                // m/z ascending (stable), as spectrum_utils orders peaks on construction
                const int64_t m = n_peaks - first_peak;
                double *mz = s->mz + first_peak;
                float *in = s->inten + first_peak;
                if (!std::is_sorted(mz, mz + m)) {
                    order.resize(m);
                    std::iota(order.begin(), order.end(), 0);
                    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t c) { return mz[a] < mz[c]; });
                    tmz.assign(mz, mz + m);
                    tin.assign(in, in + m);
                    for (int64_t i = 0; i < m; ++i) {
                        mz[i] = tmz[order[i]];
                        in[i] = tin[order[i]];
                    }
                }
            }
            n_title += ie - ib;
            if (seq_b) n_seq += seq_e - seq_b;
            ++n_spec;
            inside = false;
            continue;
        }
        if (*b == '#' || *b == ';' || *b == '!' || *b == '/') continue;
        const char *eq = (const char *)memchr(b, '=', (size_t)(e - b));
        if (eq) {
            const char *kb = b, *ke = eq, *vb = eq + 1, *ve = e;
            strip(kb, ke);
            strip(vb, ve);
            if (ieq(kb, ke, "TITLE")) {
                title_b = vb;
                title_e = ve;
            } else if (ieq(kb, ke, "SCAN") || ieq(kb, ke, "SCANS")) {
                scan_b = vb;
                scan_e = ve;
            } else if (ieq(kb, ke, "PEPMASS")) {
                SOLO_REQUIRE(first_number(vb, ve, pm), SOLO_EINVAL, "'%s' line %lld: PEPMASS is not a number", path,
                             (long long)line_no);
                have_pm = true;
            } else if (ieq(kb, ke, "CHARGE")) {
                z = parse_charge(vb, ve);
            } else if (ieq(kb, ke, "RTINSECONDS")) {
                SOLO_REQUIRE(first_number(vb, ve, rt), SOLO_EINVAL, "'%s' line %lld: RTINSECONDS is not a number", path,
                             (long long)line_no);
            } else if (ieq(kb, ke, "SEQ")) {
                seq_b = vb;
                seq_e = ve;
            } else if (ieq(kb, ke, "DECOY")) {
                decoy = true;
            }
            continue;
        }
        double m, in;
        const char *rest;
        SOLO_REQUIRE(first_number(b, e, m, &rest) && first_number(rest, e, in), SOLO_EINVAL,
                     "'%s' line %lld: expected a peak 'm/z intensity'", path, (long long)line_no);
        if (s) {
            s->mz[n_peaks] = m;
            s->inten[n_peaks] = (float)in;
        }
        ++n_peaks;
    }
    SOLO_REQUIRE(!inside, SOLO_EINVAL, "'%s': BEGIN IONS without END IONS at the end of the file", path);
    if (s) {
        s->peak_off[n_spec] = n_peaks;
        s->title_off[n_spec] = n_title;
        s->seq_off[n_spec] = n_seq;
    }
}

}  // namespace

void mgf_count(const char *path, int64_t *n_spectra, int64_t *n_peaks, int64_t *n_title_bytes, int64_t *n_seq_bytes) {
    TextFile f;
    f.open(path);
    walk_mgf(f, path, nullptr, *n_spectra, *n_peaks, *n_title_bytes, *n_seq_bytes);
}

void mgf_read(const char *path, int64_t n_spectra, int64_t n_peaks, int64_t n_title_bytes, int64_t n_seq_bytes,
              double *prec_mz, int32_t *prec_charge, double *rt, uint8_t *is_decoy, int64_t *peak_off, double *mz,
              float *inten, int64_t *title_off, char *titles, int64_t *seq_off, char *seqs) {
    TextFile f;
    f.open(path);
    int64_t a, b, c, d;
    walk_mgf(f, path, nullptr, a, b, c, d);
    SOLO_REQUIRE(a == n_spectra && b == n_peaks && c == n_title_bytes && d == n_seq_bytes, SOLO_EINVAL,
                 "'%s' changed between the counting and the reading pass", path);
    MgfSink s{prec_mz, prec_charge, rt, is_decoy, peak_off, mz, inten, title_off, titles, seq_off, seqs};
    walk_mgf(f, path, &s, a, b, c, d);
}

}  // namespace solo
