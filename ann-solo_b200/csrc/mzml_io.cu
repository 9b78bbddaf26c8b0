// mzML query files straight into peak-store (CSR) arrays.
//
// Replaces reader.read_mzml / _parse_spectrum_mzml (reference reader.py:659-741, pyteomics + lxml, one
// dict + one MsmsSpectrum per spectrum) for the query side of SURVEY.md §8f N3. The memory-mapped XML is
// scanned for <spectrum> elements; inside each, only what the reference reads is interpreted (by PSI-MS
// accession): ms level (MS:1000511), scan start time of the first scan (MS:1000016), of the first
// selected ion its m/z (MS:1000744), charge state (MS:1000041) or possible charge state (MS:1000633),
// and the two binary arrays m/z (MS:1000514) and intensity (MS:1000515), 32- or 64-bit floats
// (MS:1000521 / MS:1000523), uncompressed or zlib (MS:1000576 / MS:1000574; zlib is loaded on first
// use). Semantics kept from the reference:
//   * only MS level 2 spectra are returned; `index` is the position among ALL spectra of the file (:677)
//   * identifier = the integer after "scan=" in the spectrum id, else after "index="; a spectrum whose
//     id has neither, or where the rest of the id is not an integer, is skipped (ValueError -> warning
//     in the reference, :708-714); so is an MS2 spectrum without a selected ion
//   * precursor charge 0 = neither charge cvParam (None in the reference)
//   * retention time is the value as written (pyteomics hands the reference a unit-carrying float)
// Not handled (SOLO_EINVAL, loudly): numpress / other compressions, referenceableParamGroupRef inside
// spectra, non-float array types. pyteomics is absent: parity unpinned beyond the reader source.
#include <dlfcn.h>
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

struct XmlFile {
    const char *p = nullptr;
    size_t size = 0;
    int fd = -1;
    ~XmlFile() {
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

const char *find(const char *b, const char *e, const char *needle) {
    const size_t n = strlen(needle);
    if ((size_t)(e - b) < n) return nullptr;
    return (const char *)memmem(b, (size_t)(e - b), needle, n);
}

// value of attribute `name` inside the tag text [b, e); false when absent
bool attr(const char *b, const char *e, const char *name, const char *&vb, const char *&ve) {
    const size_t n = strlen(name);
    for (const char *p = b; (p = find(p, e, name)) != nullptr; p += n) {
        if (p > b && !(p[-1] == ' ' || p[-1] == '\t' || p[-1] == '\n' || p[-1] == '\r')) continue;
        const char *q = p + n;
        while (q < e && (*q == ' ' || *q == '\t')) ++q;
        if (q >= e || *q != '=') continue;
        ++q;
        while (q < e && (*q == ' ' || *q == '\t')) ++q;
        if (q >= e || (*q != '"' && *q != '\'')) continue;
        const char quote = *q++;
        const char *r = (const char *)memchr(q, quote, (size_t)(e - q));
        if (!r) return false;
        vb = q;
        ve = r;
        return true;
    }
    return false;
}

bool attr_is(const char *b, const char *e, const char *name, const char *value) {
    const char *vb, *ve;
    return attr(b, e, name, vb, ve) && (size_t)(ve - vb) == strlen(value) && memcmp(vb, value, strlen(value)) == 0;
}

bool to_double(const char *b, const char *e, double &v) {
    while (b < e && (*b == ' ' || *b == '+')) ++b;
    auto r = std::from_chars(b, e, v);
    return r.ec == std::errc() && r.ptr == e;
}

bool to_int(const char *b, const char *e, long long &v) {
    while (b < e && *b == ' ') ++b;
    while (e > b && e[-1] == ' ') --e;
    if (b < e && *b == '+') ++b;
    auto r = std::from_chars(b, e, v);
    return b < e && r.ec == std::errc() && r.ptr == e;
}

// RFC 4648 base64 (whitespace tolerated); returns false on a malformed string
bool base64_decode(const char *b, const char *e, std::vector<uint8_t> &out) {
    static int8_t T[256];
    static bool init = false;
    if (!init) {
        memset(T, -1, sizeof T);
        const char *abc = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
        for (int i = 0; i < 64; ++i) T[(uint8_t)abc[i]] = (int8_t)i;
        init = true;
    }
    out.clear();
    out.reserve((size_t)(e - b) / 4 * 3 + 3);
    uint32_t acc = 0;
    int bits = 0;
    for (; b < e; ++b) {
        const char c = *b;
        if (c == '=' || c == ' ' || c == '\n' || c == '\r' || c == '\t') continue;
        const int v = T[(uint8_t)c];
        if (v < 0) return false;
        acc = (acc << 6) | (uint32_t)v;
        bits += 6;
        if (bits >= 8) {
            bits -= 8;
            out.push_back((uint8_t)(acc >> bits));
        }
    }
    return true;
}

// zlib's uncompress(), resolved on first use so that the library itself has no link-time dependency
typedef int (*uncompress_fn)(unsigned char *, unsigned long *, const unsigned char *, unsigned long);
uncompress_fn zlib_uncompress() {
    static uncompress_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *h = dlopen("libz.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libz.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) fn = (uncompress_fn)dlsym(h, "uncompress");
    }
    return fn;
}

struct BinaryArray {
    const char *b = nullptr, *e = nullptr;  // base64 text
    int bits = 0;                           // 32 / 64
    bool zlib = false, is_mz = false, is_int = false, unsupported = false;
};

struct SpectrumInfo {
    long long scan_nr = 0;
    int ms_level = -1;
    bool id_ok = false, have_prec = false;
    double prec_mz = 0.0, rt = NAN;
    int charge = 0, possible_charge = 0;
    long long default_len = -1;
    BinaryArray mz, inten;
};

// decode one binary array into doubles; n_expected < 0: unknown
void decode_array(const BinaryArray &a, long long n_expected, std::vector<uint8_t> &raw, std::vector<uint8_t> &tmp,
                  std::vector<double> &out, const char *path, long long index) {
    SOLO_REQUIRE(a.bits == 32 || a.bits == 64, SOLO_EINVAL, "'%s' spectrum %lld: binary array without a float type", path,
                 index);
    SOLO_REQUIRE(base64_decode(a.b, a.e, raw), SOLO_EINVAL, "'%s' spectrum %lld: malformed base64", path, index);
    const size_t esz = (size_t)a.bits / 8;
    const uint8_t *data = raw.data();
    size_t nbytes = raw.size();
    if (a.zlib && nbytes) {
        uncompress_fn un = zlib_uncompress();
        SOLO_REQUIRE(un != nullptr, SOLO_ESTATE, "'%s' holds zlib-compressed arrays but libz.so.1 cannot be loaded", path);
        size_t cap = n_expected >= 0 ? (size_t)n_expected * esz + 64 : nbytes * 8 + 1024;
        for (int attempt = 0;; ++attempt) {
            tmp.resize(cap);
            unsigned long got = (unsigned long)cap;
            const int rc = un(tmp.data(), &got, raw.data(), (unsigned long)nbytes);
            if (rc == 0) {
                data = tmp.data();
                nbytes = got;
                break;
            }
            SOLO_REQUIRE(rc == -5 && attempt < 8, SOLO_EINVAL, "'%s' spectrum %lld: zlib error %d", path, index, rc);  // Z_BUF_ERROR
            cap *= 4;
        }
    }
    SOLO_REQUIRE(nbytes % esz == 0, SOLO_EINVAL, "'%s' spectrum %lld: %zu bytes is not a whole number of %d-bit floats",
                 path, index, nbytes, a.bits);
    const size_t n = nbytes / esz;
    out.resize(n);
    for (size_t i = 0; i < n; ++i) {
        if (a.bits == 64) {
            memcpy(&out[i], data + 8 * i, 8);
        } else {
            float f;
            memcpy(&f, data + 4 * i, 4);
            out[i] = (double)f;
        }
    }
}

// interpret one <spectrum ...> ... </spectrum> element
void parse_spectrum(const char *sb, const char *se, SpectrumInfo &s, const char *path, long long index) {
    const char *tag_end = (const char *)memchr(sb, '>', (size_t)(se - sb));
    SOLO_REQUIRE(tag_end, SOLO_EINVAL, "'%s' spectrum %lld: unterminated tag", path, index);
    const char *vb, *ve;
    if (attr(sb, tag_end, "id", vb, ve)) {
        // reader.py:708-714: int(id[id.find('scan=') + 5:]) else the same for 'index='
        const char *k = find(vb, ve, "scan=");
        size_t klen = 5;
        if (!k) {
            k = find(vb, ve, "index=");
            klen = 6;
        }
        if (k) s.id_ok = to_int(k + klen, ve, s.scan_nr);
    }
    if (attr(sb, tag_end, "defaultArrayLength", vb, ve)) to_int(vb, ve, s.default_len);
    bool in_scan = false, seen_scan = false, in_ion = false, seen_ion = false, in_bda = false;
    BinaryArray cur;
    for (const char *p = tag_end + 1; p < se;) {
        const char *lt = (const char *)memchr(p, '<', (size_t)(se - p));
        if (!lt) break;
        const char *gt = (const char *)memchr(lt, '>', (size_t)(se - lt));
        if (!gt) break;
        const bool closing = lt + 1 < gt && lt[1] == '/';
        const char *nb = lt + 1 + (closing ? 1 : 0), *ne = nb;
        while (ne < gt && *ne != ' ' && *ne != '/' && *ne != '\t' && *ne != '\n' && *ne != '\r') ++ne;
        auto is = [&](const char *w) { return (size_t)(ne - nb) == strlen(w) && memcmp(nb, w, strlen(w)) == 0; };
        p = gt + 1;
        if (closing) {
            if (is("scan")) in_scan = false;
            else if (is("selectedIon")) in_ion = false;
            else if (is("binaryDataArray")) {
                in_bda = false;
                if (cur.is_mz) s.mz = cur;
                else if (cur.is_int) s.inten = cur;
            }
            continue;
        }
        if (is("scan")) {
            in_scan = !seen_scan;   // the first scan of the scanList (:722)
            seen_scan = true;
            if (gt[-1] == '/') in_scan = false;
        } else if (is("selectedIon")) {
            in_ion = !seen_ion;     // first precursor, first selected ion (:724-725)
            if (in_ion) s.have_prec = true;
            seen_ion = true;
            if (gt[-1] == '/') in_ion = false;
        } else if (is("binaryDataArray")) {
            in_bda = true;
            cur = BinaryArray();
        } else if (is("binary")) {
            if (gt[-1] == '/') {   // <binary/>: empty array
                cur.b = cur.e = gt;
            } else {
                const char *close = find(p, se, "</binary>");
                SOLO_REQUIRE(close, SOLO_EINVAL, "'%s' spectrum %lld: <binary> without </binary>", path, index);
                cur.b = p;
                cur.e = close;
                p = close + 9;
            }
        } else if (is("referenceableParamGroupRef")) {
            SOLO_REQUIRE(false, SOLO_EINVAL, "'%s' spectrum %lld: referenceableParamGroupRef is not supported", path, index);
        } else if (is("cvParam")) {
            const char *ab, *ae;
            if (!attr(nb, gt, "accession", ab, ae)) continue;
            auto acc = [&](const char *w) { return (size_t)(ae - ab) == strlen(w) && memcmp(ab, w, strlen(w)) == 0; };
            const bool has_value = attr(nb, gt, "value", vb, ve);
            long long iv;
            if (in_bda) {
                if (acc("MS:1000523")) cur.bits = 64;
                else if (acc("MS:1000521")) cur.bits = 32;
                else if (acc("MS:1000574")) cur.zlib = true;
                else if (acc("MS:1000576")) cur.zlib = false;
                else if (acc("MS:1000514")) cur.is_mz = true;
                else if (acc("MS:1000515")) cur.is_int = true;
                else if (acc("MS:1002312") || acc("MS:1002313") || acc("MS:1002314") || acc("MS:1002746") ||
                         acc("MS:1002747") || acc("MS:1002748") || acc("MS:1000519") || acc("MS:1000522"))
                    cur.unsupported = true;  // numpress variants, integer arrays
            } else if (in_ion) {
                if (acc("MS:1000744") && has_value) to_double(vb, ve, s.prec_mz);
                else if (acc("MS:1000041") && has_value && to_int(vb, ve, iv)) s.charge = (int)iv;
                else if (acc("MS:1000633") && has_value && to_int(vb, ve, iv) && s.possible_charge == 0) s.possible_charge = (int)iv;
            } else if (in_scan) {
                if (acc("MS:1000016") && has_value) to_double(vb, ve, s.rt);
            } else if (acc("MS:1000511") && has_value && to_int(vb, ve, iv)) {
                s.ms_level = (int)iv;
            }
        }
    }
}

struct MzmlSink {  // null: counting pass
    int64_t *scan_nr;
    int32_t *index;
    double *prec_mz;
    int32_t *prec_charge;
    double *rt;
    int64_t *peak_off;
    double *mz;
    float *inten;
};

void walk_mzml(const XmlFile &f, const char *path, MzmlSink *out, int64_t &n_spec, int64_t &n_peaks, int64_t &n_skipped) {
    n_spec = n_peaks = n_skipped = 0;
    const char *p = f.p, *end = f.p + f.size;
    long long index = 0;
    std::vector<uint8_t> raw, tmp;
    std::vector<double> mz, inten;
    std::vector<int32_t> order;
    while (p && p < end) {
        const char *sb = find(p, end, "<spectrum");
        if (!sb) break;
        const char c = sb + 9 < end ? sb[9] : 0;
        if (!(c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '>')) {  // <spectrumList ...>
            p = sb + 9;
            continue;
        }
        const char *se = find(sb, end, "</spectrum>");
        SOLO_REQUIRE(se, SOLO_EINVAL, "'%s': <spectrum> %lld is not closed", path, index);
        SpectrumInfo s;
        parse_spectrum(sb, se, s, path, index);
        p = se + 11;
        const long long this_index = index++;
        if (s.ms_level != 2) continue;                     // reader.py:678
        if (!s.id_ok || !s.have_prec) {                    // ValueError / KeyError in the reference: skipped
            ++n_skipped;
            continue;
        }
        SOLO_REQUIRE(!s.mz.unsupported && !s.inten.unsupported, SOLO_EINVAL,
                     "'%s' spectrum %lld: numpress / integer binary arrays are not supported", path, this_index);
        SOLO_REQUIRE(s.mz.b && s.inten.b, SOLO_EINVAL, "'%s' spectrum %lld: m/z or intensity array missing", path, this_index);
        long long n = s.default_len;
        if (n < 0 || out) {
            decode_array(s.mz, s.default_len, raw, tmp, mz, path, this_index);
            SOLO_REQUIRE(s.default_len < 0 || (long long)mz.size() == s.default_len, SOLO_EINVAL,
                         "'%s' spectrum %lld: m/z array holds %zu values, defaultArrayLength says %lld", path, this_index,
                         mz.size(), s.default_len);
            n = (long long)mz.size();
        }
        if (out) {
            decode_array(s.inten, n, raw, tmp, inten, path, this_index);
            SOLO_REQUIRE((long long)inten.size() == n, SOLO_EINVAL, "'%s' spectrum %lld: %zu intensities for %lld m/z values",
                         path, this_index, inten.size(), n);
            out->scan_nr[n_spec] = s.scan_nr;
            out->index[n_spec] = (int32_t)this_index;
            out->prec_mz[n_spec] = s.prec_mz;
            out->prec_charge[n_spec] = s.charge != 0 ? s.charge : s.possible_charge;
            out->rt[n_spec] = s.rt;
            out->peak_off[n_spec] = n_peaks;
            order.resize((size_t)n);
            std::iota(order.begin(), order.end(), 0);
            if (!std::is_sorted(mz.begin(), mz.end()))
                std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return mz[a] < mz[b]; });
            for (long long i = 0; i < n; ++i) {
                out->mz[n_peaks + i] = mz[order[i]];
                out->inten[n_peaks + i] = (float)inten[order[i]];
            }
        }
        n_peaks += n;
        ++n_spec;
    }
    if (out) out->peak_off[n_spec] = n_peaks;
}

// ---------------------------------------------------------------- mzXML (reference reader.py:743-811)
// <scan num=".." msLevel=".." retentionTime="PT..S"> <precursorMz precursorCharge="..">m/z</precursorMz>
// <peaks precision="32|64" byteOrder="network" contentType="m/z-int" compressionType="none|zlib">base64</peaks>
// [nested <scan> children] </scan>. A scan's own precursorMz / peaks precede its children, so they are the
// ones found before the next <scan start tag. identifier = str(int(num)); index = position among all scans.
double be_float(const uint8_t *p, int bits) {
    if (bits == 64) {
        uint64_t v = 0;
        for (int i = 0; i < 8; ++i) v = (v << 8) | p[i];
        double d;
        memcpy(&d, &v, 8);
        return d;
    }
    uint32_t v = 0;
    for (int i = 0; i < 4; ++i) v = (v << 8) | p[i];
    float f;
    memcpy(&f, &v, 4);
    return (double)f;
}

void walk_mzxml(const XmlFile &f, const char *path, MzmlSink *out, int64_t &n_spec, int64_t &n_peaks, int64_t &n_skipped) {
    n_spec = n_peaks = n_skipped = 0;
    const char *end = f.p + f.size;
    long long index = 0;
    std::vector<uint8_t> raw, tmp;
    std::vector<double> mz, inten;
    std::vector<int32_t> order;
    auto next_scan = [&](const char *from) -> const char * {
        for (const char *q = from; q && (q = find(q, end, "<scan")) != nullptr; q += 5) {
            const char c = q + 5 < end ? q[5] : 0;
            if (c == ' ' || c == '\t' || c == '\n' || c == '\r') return q;
        }
        return nullptr;
    };
    for (const char *sb = f.size ? next_scan(f.p) : nullptr; sb;) {
        const char *tag_end = (const char *)memchr(sb, '>', (size_t)(end - sb));
        SOLO_REQUIRE(tag_end, SOLO_EINVAL, "'%s': unterminated <scan> tag", path);
        const char *nxt = next_scan(tag_end);
        const char *se = nxt ? nxt : end;   // this scan's own children end before the next <scan
        const long long this_index = index++;
        const char *vb, *ve;
        long long level = -1, num = 0;
        if (attr(sb, tag_end, "msLevel", vb, ve)) to_int(vb, ve, level);
        const bool id_ok = attr(sb, tag_end, "num", vb, ve) && to_int(vb, ve, num);
        double rt = NAN;
        if (attr(sb, tag_end, "retentionTime", vb, ve) && ve - vb > 3 && vb[0] == 'P' && vb[1] == 'T' && ve[-1] == 'S') {
            double sec;
            if (to_double(vb + 2, ve - 1, sec)) rt = sec / 60.0;   // xsd:duration in seconds -> minutes
        }
        sb = nxt;
        if (level != 2) continue;                                   // reader.py:761
        const char *pm = find(tag_end, se, "<precursorMz");
        const char *pk = find(tag_end, se, "<peaks");
        if (!id_ok || !pm) {
            ++n_skipped;
            continue;
        }
        const char *pm_gt = (const char *)memchr(pm, '>', (size_t)(se - pm));
        const char *pm_close = pm_gt ? find(pm_gt, se, "</precursorMz>") : nullptr;
        double prec_mz = 0.0;
        SOLO_REQUIRE(pm_close && to_double(pm_gt + 1, pm_close, prec_mz), SOLO_EINVAL,
                     "'%s' scan %lld: precursorMz is not a number", path, num);
        long long charge = 0;
        if (attr(pm, pm_gt, "precursorCharge", vb, ve)) to_int(vb, ve, charge);
        SOLO_REQUIRE(pk, SOLO_EINVAL, "'%s' scan %lld: no <peaks>", path, num);
        const char *pk_gt = (const char *)memchr(pk, '>', (size_t)(se - pk));
        SOLO_REQUIRE(pk_gt, SOLO_EINVAL, "'%s' scan %lld: unterminated <peaks>", path, num);
        long long precision = 32;
        if (attr(pk, pk_gt, "precision", vb, ve)) to_int(vb, ve, precision);
        SOLO_REQUIRE(precision == 32 || precision == 64, SOLO_EINVAL, "'%s' scan %lld: precision %lld", path, num, precision);
        const bool zl = attr_is(pk, pk_gt, "compressionType", "zlib");
        SOLO_REQUIRE(zl || !attr(pk, pk_gt, "compressionType", vb, ve) || attr_is(pk, pk_gt, "compressionType", "none"),
                     SOLO_EINVAL, "'%s' scan %lld: unsupported compressionType", path, num);
        SOLO_REQUIRE(!attr(pk, pk_gt, "byteOrder", vb, ve) || attr_is(pk, pk_gt, "byteOrder", "network"), SOLO_EINVAL,
                     "'%s' scan %lld: byteOrder must be network", path, num);
        const uint8_t *data = nullptr;
        size_t nbytes = 0;
        if (pk_gt[-1] != '/') {
            const char *pk_close = find(pk_gt, se, "</peaks>");
            SOLO_REQUIRE(pk_close, SOLO_EINVAL, "'%s' scan %lld: <peaks> without </peaks>", path, num);
            SOLO_REQUIRE(base64_decode(pk_gt + 1, pk_close, raw), SOLO_EINVAL, "'%s' scan %lld: malformed base64", path, num);
            data = raw.data();
            nbytes = raw.size();
            if (zl && nbytes) {
                uncompress_fn un = zlib_uncompress();
                SOLO_REQUIRE(un != nullptr, SOLO_ESTATE, "'%s' holds zlib-compressed peaks but libz.so.1 cannot be loaded", path);
                size_t cap = nbytes * 8 + 1024;
                for (int attempt = 0;; ++attempt) {
                    tmp.resize(cap);
                    unsigned long got = (unsigned long)cap;
                    const int rc = un(tmp.data(), &got, raw.data(), (unsigned long)nbytes);
                    if (rc == 0) {
                        data = tmp.data();
                        nbytes = got;
                        break;
                    }
                    SOLO_REQUIRE(rc == -5 && attempt < 8, SOLO_EINVAL, "'%s' scan %lld: zlib error %d", path, num, rc);
                    cap *= 4;
                }
            }
        }
        const size_t pair = (size_t)precision / 4;   // bytes per (m/z, intensity) pair
        SOLO_REQUIRE(nbytes % pair == 0, SOLO_EINVAL, "'%s' scan %lld: %zu bytes is not a whole number of peaks", path, num, nbytes);
        const long long n = (long long)(nbytes / pair);
        if (out) {
            mz.resize((size_t)n);
            inten.resize((size_t)n);
            for (long long i = 0; i < n; ++i) {
                mz[i] = be_float(data + pair * i, (int)precision);
                inten[i] = be_float(data + pair * i + pair / 2, (int)precision);
            }
            out->scan_nr[n_spec] = num;
            out->index[n_spec] = (int32_t)this_index;
            out->prec_mz[n_spec] = prec_mz;
            out->prec_charge[n_spec] = (int32_t)charge;
            out->rt[n_spec] = rt;
            out->peak_off[n_spec] = n_peaks;
            order.resize((size_t)n);
            std::iota(order.begin(), order.end(), 0);
            if (!std::is_sorted(mz.begin(), mz.end()))
                std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return mz[a] < mz[b]; });
            for (long long i = 0; i < n; ++i) {
                out->mz[n_peaks + i] = mz[order[i]];
                out->inten[n_peaks + i] = (float)inten[order[i]];
            }
        }
        n_peaks += n;
        ++n_spec;
    }
    if (out) out->peak_off[n_spec] = n_peaks;
}

}  // namespace

void mzxml_count(const char *path, int64_t *n_spectra, int64_t *n_peaks, int64_t *n_skipped) {
    XmlFile f;
    f.open(path);
    walk_mzxml(f, path, nullptr, *n_spectra, *n_peaks, *n_skipped);
}

void mzxml_read(const char *path, int64_t n_spectra, int64_t n_peaks, int64_t *scan_nr, int32_t *index, double *prec_mz,
                int32_t *prec_charge, double *rt, int64_t *peak_off, double *mz, float *inten) {
    XmlFile f;
    f.open(path);
    int64_t a, b, c;
    walk_mzxml(f, path, nullptr, a, b, c);
    SOLO_REQUIRE(a == n_spectra && b == n_peaks, SOLO_EINVAL, "'%s' holds %lld MS2 scans / %lld peaks, the buffers were sized for %lld / %lld",
                 path, (long long)a, (long long)b, (long long)n_spectra, (long long)n_peaks);
    MzmlSink s{scan_nr, index, prec_mz, prec_charge, rt, peak_off, mz, inten};
    walk_mzxml(f, path, &s, a, b, c);
}

void mzml_count(const char *path, int64_t *n_spectra, int64_t *n_peaks, int64_t *n_skipped) {
    XmlFile f;
    f.open(path);
    walk_mzml(f, path, nullptr, *n_spectra, *n_peaks, *n_skipped);
}

void mzml_read(const char *path, int64_t n_spectra, int64_t n_peaks, int64_t *scan_nr, int32_t *index, double *prec_mz,
               int32_t *prec_charge, double *rt, int64_t *peak_off, double *mz, float *inten) {
    XmlFile f;
    f.open(path);
    int64_t a, b, c;
    walk_mzml(f, path, nullptr, a, b, c);
    SOLO_REQUIRE(a == n_spectra && b == n_peaks, SOLO_EINVAL, "'%s' holds %lld MS2 spectra / %lld peaks, the buffers were sized for %lld / %lld",
                 path, (long long)a, (long long)b, (long long)n_spectra, (long long)n_peaks);
    MzmlSink s{scan_nr, index, prec_mz, prec_charge, rt, peak_off, mz, inten};
    walk_mzml(f, path, &s, a, b, c);
}

}  // namespace solo
