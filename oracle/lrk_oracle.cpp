// =====================================================================================
// oracle/lrk_oracle.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (C-style C++17, fp64, single thread unless stated) of the LibRec 3.0.0
// matrix-factorisation hot path, used ONLY as the checker by tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs.  Nothing under librec_b200/ may
// import, link or call it.
//
// PARITY UNPINNED: the reference (pure Java) cannot be executed in the build container
// (no JVM) and its own tests for this path assert nothing, so there are
// no golden vectors from the reference itself.  What IS pinned (tests/test_oracle_*.py):
//   * java.util.Random / nextGaussian / fdlibm log  against published JDK-8 known answers,
//   * the loader against TextDataModelTestCase's row counts (matrix4by4.txt -> 13 entries),
//   * the splitter against RatioDataSplitterTestCase's |ratio-0.8| <= 0.01 bound,
//   * hand-worked one-update micro cases for BiasedMF / PMF / BPR,
//   * java.util.PriorityQueue heap order against an independent pure-Python replay.
//
// Every function cites the reference file:line it restates (paths relative to
// /root/reference/core/src/main/java/net/librec/).  JDK behaviour (not vendored by the
// reference; JDK 8, pom.xml:17) is restated from its published algorithms, see SURVEY.md section 9.
// Build: oracle/Makefile  (g++ -O2 -ffp-contract=off: Java never contracts a*b+c into an FMA).
// =====================================================================================
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <dirent.h>
#include <sys/stat.h>
#include <string>
#include <vector>
#include <unordered_map>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LRO_API extern "C" __attribute__((visibility("default")))

// -------------------------------------------------------------------------------------
// fdlibm __ieee754_log  ==  java.lang.StrictMath.log  (called by Random.nextGaussian).
// Restated from the published fdlibm 5.3 e_log.c algorithm (Sun Microsystems, freely
// distributable) so that Gaussian init is bit-identical to a JVM, independent of glibc.
// -------------------------------------------------------------------------------------
static inline int32_t hi_word(double x) { uint64_t b; memcpy(&b, &x, 8); return (int32_t)(b >> 32); }
static inline uint32_t lo_word(double x) { uint64_t b; memcpy(&b, &x, 8); return (uint32_t)b; }
static inline double with_hi(double x, int32_t hi) {
    uint64_t b; memcpy(&b, &x, 8);
    b = (b & 0xffffffffULL) | ((uint64_t)(uint32_t)hi << 32);
    memcpy(&x, &b, 8); return x;
}

static double fdlibm_log(double x) {
    static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                        two54 = 1.80143985094819840000e+16,
                        Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
                        Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                        Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                        Lg7 = 1.479819860511658591e-01;
    const double zero = 0.0;
    double hfsq, f, s, z, R, w, t1, t2, dk;
    int32_t k, hx, i, j;
    uint32_t lx;
    hx = hi_word(x); lx = lo_word(x);
    k = 0;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return -two54 / zero;
        if (hx < 0) return (x - x) / zero;
        k -= 54; x *= two54; hx = hi_word(x);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    i = (hx + 0x95f64) & 0x100000;
    x = with_hi(x, hx | (i ^ 0x3ff00000));
    k += (i >> 20);
    f = x - 1.0;
    if ((0x000fffff & (2 + hx)) < 3) {
        if (f == zero) { if (k == 0) return zero; dk = (double)k; return dk * ln2_hi + dk * ln2_lo; }
        R = f * f * (0.5 - 0.33333333333333333 * f);
        if (k == 0) return f - R;
        dk = (double)k; return dk * ln2_hi - ((R - dk * ln2_lo) - f);
    }
    s = f / (2.0 + f);
    dk = (double)k;
    z = s * s;
    i = hx - 0x6147a;
    w = z * z;
    j = 0x6b851 - hx;
    t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    R = t2 + t1;
    if (i > 0) {
        hfsq = 0.5 * f * f;
        if (k == 0) return f - (hfsq - s * (hfsq + R));
        return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
    } else {
        if (k == 0) return f - s * (f - R);
        return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
    }
}
LRO_API double lro_strictmath_log(double x) { return fdlibm_log(x); }

// -------------------------------------------------------------------------------------
// java.util.Random (JDK 8) -- the ONE global generator behind math/algorithm/Randoms.java:31
// (static Random r), seeded once by job/RecommenderJob.java:72-79.
// -------------------------------------------------------------------------------------
struct JRandom {
    uint64_t seed = 0;
    bool have_next = false;
    double next_gauss = 0.0;
};
static JRandom g_rng;

static inline void jr_seed(JRandom* r, int64_t s) {
    r->seed = ((uint64_t)s ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1);
    r->have_next = false;
}
static inline int32_t jr_next(JRandom* r, int bits) {
    r->seed = (r->seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
    return (int32_t)(uint32_t)(r->seed >> (48 - bits));
}
static inline int32_t jr_next_int(JRandom* r) { return jr_next(r, 32); }
static inline int32_t jr_next_int_bound(JRandom* r, int32_t bound) {
    int32_t v = jr_next(r, 31);
    int32_t m = bound - 1;
    if ((bound & m) == 0) {
        v = (int32_t)(((int64_t)bound * (int64_t)v) >> 31);
    } else {
        int32_t u = v;
        for (;;) {
            v = u % bound;
            // Java int arithmetic wraps: "u - r + m < 0"
            int32_t t = (int32_t)((uint32_t)u - (uint32_t)v + (uint32_t)m);
            if (t >= 0) break;
            u = jr_next(r, 31);
        }
    }
    return v;
}
static inline double jr_next_double(JRandom* r) {
    int64_t a = jr_next(r, 26);
    int64_t b = jr_next(r, 27);
    return (double)((a << 27) + b) * 0x1.0p-53;
}
static inline double jr_next_gaussian(JRandom* r) {
    if (r->have_next) { r->have_next = false; return r->next_gauss; }
    double v1, v2, s;
    do {
        v1 = 2 * jr_next_double(r) - 1;
        v2 = 2 * jr_next_double(r) - 1;
        s = v1 * v1 + v2 * v2;
    } while (s >= 1 || s == 0);
    double multiplier = sqrt(-2 * fdlibm_log(s) / s);  // StrictMath.sqrt is IEEE-exact
    r->next_gauss = v2 * multiplier;
    r->have_next = true;
    return v1 * multiplier;
}

// Randoms.seed / uniform(int) / uniform() / gaussian : math/algorithm/Randoms.java:41-58,117-130,158-160
LRO_API void lro_seed(int64_t seed) { jr_seed(&g_rng, seed); }
LRO_API int32_t lro_next_int() { return jr_next_int(&g_rng); }
LRO_API int32_t lro_uniform_int(int32_t range) { return 0 + jr_next_int_bound(&g_rng, range - 0); }
LRO_API double lro_uniform() { return 0.0 + (1.0 - 0.0) * jr_next_double(&g_rng); }
LRO_API double lro_next_gaussian() { return jr_next_gaussian(&g_rng); }
LRO_API double lro_gaussian(double mu, double sigma) { return mu + sigma * jr_next_gaussian(&g_rng); }
// DenseMatrix.init -> assign row-major (math/structure/DenseMatrix.java:77-97); DenseVector.init (DenseVector.java:26-28)
LRO_API void lro_gaussian_fill(double* out, int64_t n, double mu, double sigma) {
    for (int64_t i = 0; i < n; ++i) out[i] = lro_gaussian(mu, sigma);
}
LRO_API void lro_rng_get_state(uint64_t* seed, int32_t* have, double* nextg) {
    *seed = g_rng.seed; *have = g_rng.have_next; *nextg = g_rng.next_gauss;
}
LRO_API void lro_rng_set_state(uint64_t seed, int32_t have, double nextg) {
    g_rng.seed = seed; g_rng.have_next = have != 0; g_rng.next_gauss = nextg;
}

// Float.valueOf(str) then widening to double (conf/Configuration.java:232-239; the float fields
// learnRate/regUser/regItem at recommender/MatrixFactorizationRecommender.java:16,54,59).
LRO_API double lro_float_promote(const char* s) { return (double)strtof(s, nullptr); }

// -------------------------------------------------------------------------------------
// Loader: data/convertor/TextDataConvertor.java:142-200 (regex "[\t;, ]" from
// data/model/TextDataModel.java:65; first blank line stops the file :176-178),
// math/structure/DataFrame.java:370-379 (first-seen dense ids), :237-261 (reverse scan into a
// table => on a duplicate (u,i) the EARLIEST line wins; binThold>=0 -> rate>thold ? 1.0 : -1.0),
// rows sorted by column (math/structure/VectorBasedSequentialSparseVector.java:74-117).
// -------------------------------------------------------------------------------------
struct LroCsr {
    int32_t U = 0, I = 0;
    std::vector<int64_t> rowptr;
    std::vector<int32_t> col;
    std::vector<double> val;
    std::vector<int64_t> date;            // UIRT only: Long.parseLong of the 4th column of the line that won (DataFrame.java:112-113, 262-270)
    std::vector<std::string> user_ids, item_ids;
};

static void split_fields(const std::string& line, const char* seps, std::vector<std::string>& out) {
    out.clear();
    std::string cur;
    for (char c : line) {
        if (strchr(seps, c) != nullptr && c != '\0') { out.push_back(cur); cur.clear(); }
        else cur.push_back(c);
    }
    out.push_back(cur);
    while (!out.empty() && out.back().empty()) out.pop_back();  // Pattern.split drops trailing empties
}
static bool is_blank(const std::string& s) {
    for (char c : s) if ((unsigned char)c > ' ') return false;  // String.trim(): chars <= U+0020
    return true;
}

// data/convertor/TextDataConvertor.java:142-200 with data/model/TextDataModel.java:58-64: `paths` is the ':'-separated
// data.input.path (already prefixed with dfs.data.dir); a directory contributes every file below it (Files.walkFileTree;
// the listing is sorted by name here, the JDK takes the file system's order); `fmt` "UIRT" requires a 4th (date) column,
// which the preference matrix ignores; the first blank line ends the CURRENT file only.
static void lro_collect_files(const std::string& path, std::vector<std::string>& out) {
    struct stat st;
    if (stat(path.c_str(), &st) != 0) return;
    if (!S_ISDIR(st.st_mode)) { out.push_back(path); return; }
    DIR* d = opendir(path.c_str());
    if (!d) return;
    std::vector<std::string> names;
    while (struct dirent* e = readdir(d)) {
        const std::string n = e->d_name;
        if (n != "." && n != "..") names.push_back(n);
    }
    closedir(d);
    std::sort(names.begin(), names.end());
    for (const std::string& n : names) lro_collect_files(path + "/" + n, out);
}

struct LroIds {                       // DataFrame's STATIC id maps (DataFrame.java:48,370-379): shared by every convertor of a run
    std::unordered_map<std::string, int32_t> umap, imap;
    std::vector<std::string> user_ids, item_ids;
};
struct LroLines { std::vector<int32_t> us, is; std::vector<double> rs; std::vector<int64_t> ds; };

static bool lro_read_lines(const char* paths, size_t need, LroIds& ids, LroLines& L) {
    std::vector<std::string> files;
    {
        const std::string all(paths);
        size_t from = 0;
        for (;;) {
            const size_t c = all.find(':', from);
            lro_collect_files(all.substr(from, c == std::string::npos ? std::string::npos : c - from), files);
            if (c == std::string::npos) break;
            from = c + 1;
        }
    }
    if (files.empty()) return false;
    std::vector<std::string> f;
    std::string line;
    static char buf[1 << 16];
    for (const std::string& path : files) {
        FILE* fp = fopen(path.c_str(), "rb");
        if (!fp) return false;
        while (fgets(buf, sizeof buf, fp)) {
            line.assign(buf);
            while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
            if (is_blank(line)) break;
            split_fields(line, "\t;, ", f);
            if (f.size() < need) continue;
            auto iu = ids.umap.find(f[0]);
            int32_t u;
            if (iu == ids.umap.end()) { u = (int32_t)ids.umap.size(); ids.umap.emplace(f[0], u); ids.user_ids.push_back(f[0]); }
            else u = iu->second;
            auto ii = ids.imap.find(f[1]);
            int32_t i;
            if (ii == ids.imap.end()) { i = (int32_t)ids.imap.size(); ids.imap.emplace(f[1], i); ids.item_ids.push_back(f[1]); }
            else i = ii->second;
            L.us.push_back(u); L.is.push_back(i); L.rs.push_back(strtod(f[2].c_str(), nullptr));
            L.ds.push_back(need == 4 ? strtoll(f[3].c_str(), nullptr, 10) : 0);
        }
        fclose(fp);
    }
    return true;
}

// DataFrame.toSparseMatrix (DataFrame.java:237-261): earliest line wins, rows sorted by item, optional binarisation
static LroCsr* lro_build_csr(const LroIds& ids, const LroLines& L, size_t need, double bin_thold) {
    LroCsr* m = new LroCsr();
    m->U = (int32_t)ids.umap.size(); m->I = (int32_t)ids.imap.size();
    m->user_ids = ids.user_ids; m->item_ids = ids.item_ids;
    const size_t n = L.us.size();
    std::vector<size_t> ord(n);
    for (size_t t = 0; t < n; ++t) ord[t] = t;
    std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) {
        if (L.us[a] != L.us[b]) return L.us[a] < L.us[b];
        return L.is[a] < L.is[b];
    });
    m->rowptr.assign((size_t)m->U + 1, 0);
    for (size_t t = 0; t < n; ++t) {
        size_t a = ord[t];
        if (t > 0 && L.us[ord[t - 1]] == L.us[a] && L.is[ord[t - 1]] == L.is[a]) continue;
        double r = L.rs[a];
        if (bin_thold >= 0) r = r > bin_thold ? 1.0 : -1.0;
        m->col.push_back(L.is[a]); m->val.push_back(r);
        if (need == 4) m->date.push_back(L.ds[a]);
        m->rowptr[(size_t)L.us[a] + 1]++;
    }
    for (int32_t u = 0; u < m->U; ++u) m->rowptr[u + 1] += m->rowptr[u];
    return m;
}

LRO_API void* lro_csr_load_paths(const char* paths, const char* fmt, double bin_thold) {
    const size_t need = (fmt && (strcmp(fmt, "UIRT") == 0 || strcmp(fmt, "uirt") == 0)) ? 4 : 3;
    LroIds ids; LroLines L;
    if (!lro_read_lines(paths, need, ids, L)) return nullptr;
    return lro_build_csr(ids, L, need, bin_thold);
}

// data/splitter/GivenTestSetDataSplitter.java:64-97 (data.model.splitter=testset): the test file(s) go through a second
// TextDataConvertor that CONTINUES the static id maps, both matrices are then built with the final dimensions, and the train
// matrix is the preference matrix with every (user, item) of the test matrix set to zero and reshaped away.
// *pref_out (optional), *train_out, *test_out are LroCsr handles to free with lro_csr_free.  returns 0 on an unreadable path.
LRO_API int32_t lro_csr_load_testset(const char* paths, const char* test_paths, const char* fmt, double bin_thold,
                                     void** pref_out, void** train_out, void** test_out) {
    const size_t need = (fmt && (strcmp(fmt, "UIRT") == 0 || strcmp(fmt, "uirt") == 0)) ? 4 : 3;
    LroIds ids; LroLines L, T;
    if (!lro_read_lines(paths, need, ids, L) || !lro_read_lines(test_paths, need, ids, T)) return 0;
    LroCsr* pref = lro_build_csr(ids, L, need, bin_thold);
    LroCsr* test = lro_build_csr(ids, T, need, bin_thold);
    LroCsr* train = new LroCsr();
    train->U = pref->U; train->I = pref->I; train->user_ids = pref->user_ids; train->item_ids = pref->item_ids;
    train->rowptr.assign((size_t)pref->U + 1, 0);
    for (int32_t u = 0; u < pref->U; ++u) {
        const int32_t* tb = test->col.data() + test->rowptr[u];
        const int32_t* te = test->col.data() + test->rowptr[u + 1];
        for (int64_t e = pref->rowptr[u]; e < pref->rowptr[u + 1]; ++e) {
            if (std::binary_search(tb, te, pref->col[(size_t)e]) || pref->val[(size_t)e] == 0.0) continue;
            train->col.push_back(pref->col[(size_t)e]); train->val.push_back(pref->val[(size_t)e]);
            if (need == 4) train->date.push_back(pref->date[(size_t)e]);
        }
        train->rowptr[(size_t)u + 1] = (int64_t)train->col.size();
    }
    if (pref_out) *pref_out = pref; else delete pref;
    *train_out = train; *test_out = test;
    return 1;
}
LRO_API void* lro_csr_load_text(const char* path, double bin_thold) { return lro_csr_load_paths(path, "UIR", bin_thold); }
LRO_API void lro_csr_dims(void* h, int32_t* U, int32_t* I, int64_t* nnz) {
    LroCsr* m = (LroCsr*)h; *U = m->U; *I = m->I; *nnz = (int64_t)m->col.size();
}
LRO_API void lro_csr_copy(void* h, int64_t* rowptr, int32_t* col, double* val) {
    LroCsr* m = (LroCsr*)h;
    memcpy(rowptr, m->rowptr.data(), m->rowptr.size() * 8);
    memcpy(col, m->col.data(), m->col.size() * 4);
    memcpy(val, m->val.data(), m->val.size() * 8);
}
// datetime matrix values aligned with the entries (UIRT loads only); returns 0 when the load had no date column
LRO_API int32_t lro_csr_copy_dates(void* h, int64_t* out) {
    LroCsr* m = (LroCsr*)h;
    if (m->date.size() != m->col.size()) return 0;
    memcpy(out, m->date.data(), m->date.size() * 8);
    return 1;
}
// outer (raw string) id of an inner user/item index, parsed as integer (-1 when not numeric)
LRO_API int64_t lro_csr_outer_id(void* h, int32_t is_item, int32_t inner) {
    LroCsr* m = (LroCsr*)h;
    const std::string& s = is_item ? m->item_ids[inner] : m->user_ids[inner];
    char* e = nullptr; long long v = strtoll(s.c_str(), &e, 10);
    return (e && *e == 0) ? (int64_t)v : -1;
}
LRO_API void lro_csr_free(void* h) { delete (LroCsr*)h; }

// -------------------------------------------------------------------------------------
// Splitter: data/splitter/RatioDataSplitter.java:136-156 -- one Randoms.uniform() per entry of
// the preference matrix in CSR order; rdm < ratio -> train.  reshape() then drops exact 0.0
// values from both sides (math/structure/OrderedIntDoubleMapping.java:342-359).
// is_train[e] = 1 train, 0 test, 2 dropped (value == 0.0).
// -------------------------------------------------------------------------------------
LRO_API void lro_split_ratio(int64_t nnz, const double* val, double ratio, uint8_t* is_train) {
    for (int64_t e = 0; e < nnz; ++e) {
        double rdm = lro_uniform();
        uint8_t t = rdm < ratio ? 1 : 0;
        if (val[e] == 0.0) t = 2;
        is_train[e] = t;
    }
}

// -------------------------------------------------------------------------------------
// The other splitters of data/splitter/ (SURVEY.md 8f, row N2 widened).  All work on the stored entries of the
// preference matrix in CSR order and consume the global java.util.Random exactly like the reference; the outputs are
// per-entry assignments (the caller drops exact zeros afterwards, like reshape()).  The *date variants need the
// datetime matrix and the three-way "valid" split a validation set: not restated.
// -------------------------------------------------------------------------------------
// entry order of a column walk (SequentialAccessSparseMatrix.column(j): rows ascending): csc[t] = CSR entry index
static void lro_csc_order(int32_t U, int32_t I, const int64_t* rowptr, const int32_t* col, std::vector<int64_t>& colptr,
                          std::vector<int64_t>& csc) {
    const int64_t nnz = rowptr[U];
    colptr.assign((size_t)I + 1, 0);
    for (int64_t e = 0; e < nnz; ++e) colptr[(size_t)col[e] + 1]++;
    for (int32_t j = 0; j < I; ++j) colptr[(size_t)j + 1] += colptr[(size_t)j];
    csc.assign((size_t)nnz, 0);
    std::vector<int64_t> fill(colptr.begin(), colptr.end() - 1);
    for (int32_t u = 0; u < U; ++u)
        for (int64_t e = rowptr[u]; e < rowptr[u + 1]; ++e) csc[(size_t)fill[(size_t)col[e]]++] = e;
}

// RatioDataSplitter.getRatioByItem (RatioDataSplitter.java:315-334): one Randoms.uniform() per entry in COLUMN order;
// getRatioByUser (:232-249) draws in row order and is identical to getRatioByRating (lro_split_ratio).
LRO_API void lro_split_ratio_item(int32_t U, int32_t I, const int64_t* rowptr, const int32_t* col, double ratio, uint8_t* is_train) {
    std::vector<int64_t> colptr, csc;
    lro_csc_order(U, I, rowptr, col, colptr, csc);
    for (size_t t = 0; t < csc.size(); ++t) is_train[csc[t]] = lro_uniform() < ratio ? 1 : 0;
}

// RatioDataSplitter.getRatio(trainRatio, validationRatio) (RatioDataSplitter.java:382-412): one Randoms.uniform() per entry in
// CSR order; < trainRatio -> train (0), < trainRatio + validationRatio -> validation (1), else test (2).  Nothing is split
// unless both ratios are positive and their sum is below 1 (assign stays 0 = everything "train"; the reference then leaves
// all three matrices null).
LRO_API void lro_split_ratio_valid(int64_t nnz, double train_ratio, double valid_ratio, uint8_t* assign) {
    for (int64_t e = 0; e < nnz; ++e) assign[e] = 0;
    if (!((train_ratio > 0 && valid_ratio > 0) && (train_ratio + valid_ratio) < 1)) return;
    for (int64_t e = 0; e < nnz; ++e) {
        const double rdm = lro_uniform();
        assign[e] = rdm < train_ratio ? 0 : (rdm < train_ratio + valid_ratio ? 1 : 2);
    }
}

// KCVDataSplitter.splitData(kFold) (KCVDataSplitter.java:84-123): entry index -> fold key (int)(index / (numRates / numFold)) + 1,
// paired with one Randoms.uniform() each, sorted DESCENDING by the random value (Lists.sortList(.., true), stable); the
// entries in CSR order receive the keys in that sorted order.  fold_out[e] in 1..numFold; fold k's test set = {fold == k}.
LRO_API void lro_split_kcv(int64_t nnz, int32_t k_fold, int32_t* fold_out) {
    const int64_t num_fold = k_fold > nnz ? nnz : k_fold;
    const double indv = ((double)nnz + 0.0) / (double)num_fold;
    std::vector<std::pair<int32_t, double>> rdm((size_t)nnz);
    for (int64_t i = 0; i < nnz; ++i) rdm[(size_t)i] = {(int32_t)((double)i / indv) + 1, lro_uniform()};
    std::stable_sort(rdm.begin(), rdm.end(), [](const std::pair<int32_t, double>& a, const std::pair<int32_t, double>& b) { return a.second > b.second; });
    for (int64_t i = 0; i < nnz; ++i) fold_out[i] = rdm[(size_t)i].first;
}

// LOOCVDataSplitter.getLOOByUser / getLOOByItems (LOOCVDataSplitter.java:144-165, 197-216): one entry per non-empty row
// (column) at position (int)(n * Randoms.uniform()) goes to the test set.
LRO_API void lro_split_loocv(int32_t U, int32_t I, const int64_t* rowptr, const int32_t* col, int32_t by_item, uint8_t* is_train) {
    const int64_t nnz = rowptr[U];
    for (int64_t e = 0; e < nnz; ++e) is_train[e] = 1;
    if (!by_item) {
        for (int32_t u = 0; u < U; ++u) {
            const int64_t n = rowptr[u + 1] - rowptr[u];
            if (n == 0) continue;
            is_train[rowptr[u] + (int64_t)((double)n * lro_uniform())] = 0;
        }
    } else {
        std::vector<int64_t> colptr, csc;
        lro_csc_order(U, I, rowptr, col, colptr, csc);
        for (int32_t j = 0; j < I; ++j) {
            const int64_t n = colptr[(size_t)j + 1] - colptr[(size_t)j];
            if (n == 0) continue;
            is_train[csc[(size_t)(colptr[(size_t)j] + (int64_t)((double)n * lro_uniform()))]] = 0;
        }
    }
}

// Randoms.nextIntArray(length, range) (math/algorithm/Randoms.java:576-604, nextInt :520-535): `length` distinct draws of
// r.nextInt(range), repeats rejected, sorted ascending; the identity when range == length.
static void lro_next_int_array(int32_t length, int32_t range, std::vector<int32_t>& out) {
    out.clear();
    if (range == length) { for (int32_t i = 0; i < length; ++i) out.push_back(i); return; }
    while ((int32_t)out.size() < length) {
        const int32_t next = lro_uniform_int(range);
        if (std::find(out.begin(), out.end(), next) == out.end()) out.push_back(next);
    }
    std::sort(out.begin(), out.end());
}

// GivenNDataSplitter.getGivenNByUser / getGivenNByItem (GivenNDataSplitter.java:137-167, 217-245): a row (column) with more
// than N entries keeps N random positions for training and sends the rest to the test set; one with at most N keeps all.
LRO_API void lro_split_givenn(int32_t U, int32_t I, const int64_t* rowptr, const int32_t* col, int32_t by_item, int32_t n_given,
                              uint8_t* is_train) {
    const int64_t nnz = rowptr[U];
    for (int64_t e = 0; e < nnz; ++e) is_train[e] = 1;
    if (n_given <= 0) return;
    std::vector<int32_t> given;
    std::vector<int64_t> colptr, csc;
    if (by_item) lro_csc_order(U, I, rowptr, col, colptr, csc);
    const int32_t lines = by_item ? I : U;
    for (int32_t l = 0; l < lines; ++l) {
        const int64_t b = by_item ? colptr[(size_t)l] : rowptr[l];
        const int64_t n = (by_item ? colptr[(size_t)l + 1] : rowptr[l + 1]) - b;
        if (n <= n_given) continue;
        lro_next_int_array(n_given, (int32_t)n, given);
        size_t g = 0;
        for (int64_t pos = 0; pos < n; ++pos) {
            const int64_t e = by_item ? csc[(size_t)(b + pos)] : b + pos;
            if (g < given.size() && given[g] == pos) ++g; else is_train[e] = 0;
        }
    }
}

// Date-ordered variants (util/RatingContext.java:35-44 orders by timestamp, Collections.sort is stable).  mode:
//   0 ratio ratingdate (RatioDataSplitter.java:190-221)  all entries by date, first (int)(n * ratio) -> train
//   1 ratio userdate   (:283-313)   per user, first (int)(n_u * ratio) -> train.  QUIRK kept: the timestamp is
//                                    (long) RATING of the entry, not its date (the loop walks preferenceMatrix.row(u))
//   2 ratio itemdate   (:339-373)   per item by date
//   3 loocv userdate   (LOOCVDataSplitter.java:171-191)  per user, the latest entry -> test
//   4 loocv itemdate   (:222-250)
//   5 givenn userdate  (GivenNDataSplitter.java:176-207) per user, first N -> train; same (long) rating QUIRK as mode 1
//   6 givenn itemdate  (:254-284)   per item by date
LRO_API void lro_split_by_date(int32_t U, int32_t I, const int64_t* rowptr, const int32_t* col, const double* val,
                               const int64_t* date, int32_t mode, double ratio, int32_t n_given, uint8_t* is_train) {
    const int64_t nnz = rowptr[U];
    std::vector<int64_t> colptr, csc;
    const bool by_col = mode == 2 || mode == 4 || mode == 6;
    if (by_col) lro_csc_order(U, I, rowptr, col, colptr, csc);
    auto cut = [&](const std::vector<int64_t>& entries) {
        const size_t n = entries.size();
        if (n == 0) return;
        std::vector<std::pair<int64_t, int64_t>> ctx(n);                 // (timestamp, entry)
        for (size_t t = 0; t < n; ++t) {
            const int64_t e = entries[t];
            ctx[t] = {(mode == 1 || mode == 5) ? (int64_t)val[e] : date[e], e};
        }
        std::stable_sort(ctx.begin(), ctx.end(), [](const std::pair<int64_t, int64_t>& a, const std::pair<int64_t, int64_t>& b) {
            return (double)a.first - (double)b.first < 0;                   // compareTo works on the double difference
        });
        size_t n_train;
        if (mode <= 2) n_train = (size_t)(int32_t)((double)n * ratio);
        else if (mode <= 4) n_train = n - 1;
        else n_train = (size_t)std::min<int64_t>((int64_t)n, n_given);
        for (size_t t = 0; t < n; ++t) is_train[ctx[t].second] = t < n_train ? 1 : 0;
    };
    std::vector<int64_t> line;
    if (mode == 0) {
        line.resize((size_t)nnz);
        for (int64_t e = 0; e < nnz; ++e) line[(size_t)e] = e;
        cut(line);
    } else if (!by_col) {
        for (int32_t u = 0; u < U; ++u) {
            line.clear();
            for (int64_t e = rowptr[u]; e < rowptr[u + 1]; ++e) line.push_back(e);
            cut(line);
        }
    } else {
        for (int32_t j = 0; j < I; ++j) {
            line.assign(csc.begin() + colptr[(size_t)j], csc.begin() + colptr[(size_t)j + 1]);
            cut(line);
        }
    }
}

// -------------------------------------------------------------------------------------
// MatrixRecommender.setup: recommender/MatrixRecommender.java:88-128
// globalMean = sum in CSR order / nnz (math/structure/RowSequentialAccessSparseMatrix.java:161-167);
// minRate/maxRate from the rating set, minRate=0 if equal (:103-107).
// -------------------------------------------------------------------------------------
LRO_API void lro_matrix_setup(int64_t nnz, const double* val, double* global_mean, double* min_rate, double* max_rate) {
    double s = 0.0, mn = INFINITY, mx = -INFINITY;
    for (int64_t e = 0; e < nnz; ++e) { s += val[e]; mn = std::min(mn, val[e]); mx = std::max(mx, val[e]); }
    *global_mean = s / (double)nnz;
    if (mn == mx) mn = 0;
    *min_rate = mn; *max_rate = mx;
}

// MatrixFactorizationRecommender.setup: recommender/MatrixFactorizationRecommender.java:80-93.
// Gaussian init N(0, (double)0.001f) in the order userFactors, itemFactors, impUserFactors,
// impItemFactors (the fork's two extra matrices consume RNG draws), then for BiasedMF the
// biases userBiases, itemBiases (recommender/cf/rating/BiasedMFRecommender.java:59-63).
LRO_API void lro_mf_setup(int32_t U, int32_t I, int32_t k, double* P, double* Q, double* bu, double* bi) {
    const double mean = (double)0.0f, sd = (double)0.001f;
    lro_gaussian_fill(P, (int64_t)U * k, mean, sd);
    lro_gaussian_fill(Q, (int64_t)I * k, mean, sd);
    for (int64_t t = 0; t < (int64_t)U * k; ++t) (void)lro_gaussian(mean, sd);  // impUserFactors
    for (int64_t t = 0; t < (int64_t)I * k; ++t) (void)lro_gaussian(mean, sd);  // impItemFactors
    if (bu) lro_gaussian_fill(bu, U, mean, sd);
    if (bi) lro_gaussian_fill(bi, I, mean, sd);
}

// DenseVector.dot: math/structure/DenseVector.java:104-111 -- left to right from 0.0
static inline double dot_lr(const double* a, const double* b, int k) {
    double r = 0.0;
    for (int f = 0; f < k; ++f) r += b[f] * a[f];
    return r;
}

enum { LRO_BIASEDMF = 0, LRO_PMF = 1, LRO_BPR = 2, LRO_RANKSGD = 3 };

// predict(): BiasedMFRecommender.java:118-120 ; MatrixFactorizationRecommender.java:104-106
static inline double predict_raw(int model, int k, const double* P, const double* Q, const double* bu,
                                 const double* bi, double mu, int32_t u, int32_t i) {
    double d = dot_lr(P + (int64_t)u * k, Q + (int64_t)i * k, k);
    if (model == LRO_BIASEDMF) return d + bu[u] + bi[i] + mu;
    return d;
}

// -------------------------------------------------------------------------------------
// One BiasedMF epoch: recommender/cf/rating/BiasedMFRecommender.java:68-100.
// `order` (optional) replaces CSR order by an explicit entry permutation -- that is NOT the
// reference's behaviour; it exists to separate ordering error from kernel error.
// lr/regU/regI are the float fields promoted to double; regB is a double (:36,:56).
// -------------------------------------------------------------------------------------
LRO_API double lro_biasedmf_epoch(int32_t U, const int64_t* rowptr, const int32_t* col, const double* val,
                                  int32_t k, double* P, double* Q, double* bu, double* bi, double mu,
                                  float lr_f, float regU_f, float regI_f, double regB,
                                  const int64_t* order, const int32_t* row_of) {
    const double learnRate = (double)lr_f, regUser = (double)regU_f, regItem = (double)regI_f;
    double loss = 0.0;
    const int64_t nnz = rowptr[U];
    int32_t u = 0;
    for (int64_t t = 0; t < nnz; ++t) {
        int64_t e = order ? order[t] : t;
        if (order) u = row_of[e];
        else while (rowptr[u + 1] <= e) ++u;
        const int32_t i = col[e];
        double* pu = P + (int64_t)u * k;
        double* qi = Q + (int64_t)i * k;
        const double predictRating = dot_lr(pu, qi, k) + bu[u] + bi[i] + mu;
        const double error = val[e] - predictRating;
        loss += error * error;
        const double ub = bu[u];
        bu[u] += learnRate * (error - regB * ub);
        loss += regB * ub * ub;
        const double ib = bi[i];
        bi[i] += learnRate * (error - regB * ib);
        loss += regB * ib * ib;
        for (int f = 0; f < k; ++f) {
            const double uf = pu[f], itf = qi[f];
            pu[f] += learnRate * (error * itf - regUser * uf);
            qi[f] += learnRate * (error * uf - regItem * itf);
            loss += regUser * uf * uf + regItem * itf * itf;
        }
    }
    loss *= 0.5;
    return loss;
}

// Vanilla PMF epoch: recommender/cf/rating/PMFSimilarityRecommender.java:59-90 (identical loop kept
// commented in PMFRatingRecommender.java:130-162); predict = MatrixFactorizationRecommender.java:104-106.
LRO_API double lro_pmf_epoch(int32_t U, const int64_t* rowptr, const int32_t* col, const double* val,
                             int32_t k, double* P, double* Q, float lr_f, float regU_f, float regI_f,
                             const int64_t* order, const int32_t* row_of) {
    const double learnRate = (double)lr_f, regUser = (double)regU_f, regItem = (double)regI_f;
    double loss = 0.0;
    const int64_t nnz = rowptr[U];
    int32_t u = 0;
    for (int64_t t = 0; t < nnz; ++t) {
        int64_t e = order ? order[t] : t;
        if (order) u = row_of[e];
        else while (rowptr[u + 1] <= e) ++u;
        const int32_t i = col[e];
        double* pu = P + (int64_t)u * k;
        double* qi = Q + (int64_t)i * k;
        const double error = val[e] - dot_lr(pu, qi, k);
        loss += error * error;
        for (int f = 0; f < k; ++f) {
            const double uf = pu[f], itf = qi[f];
            pu[f] += learnRate * (error * itf - regUser * uf);
            qi[f] += learnRate * (error * uf - regItem * itf);
            loss += regUser * uf * uf + regItem * itf * itf;
        }
    }
    loss *= 0.5;
    return loss;
}

// -------------------------------------------------------------------------------------
// SVD++ (SURVEY.md 8f, row N3 -- groundwork: oracle only, no device kernel yet):
// recommender/cf/rating/SVDPlusPlusRecommender.java:62-123 (extends BiasedMF; impItemFactors Y is I x k, Gaussian-initialised
// AFTER the BiasedMF setup, regImpItem = rec.impItem.regularization, default 0.015).  Per user with a non-empty row:
// fv = |N(u)|^-1/2 * sum_{j in N(u)} y_j, FIXED for the whole row; per rating in item order: predict =
// (b_u + b_i + mu) + sum_f (fv_f + p_uf) q_if (:126-135), the bias and factor updates of BiasedMF except that the item
// gradient uses (p_uf + fv_f) (:93) and steps_f += e * q_if(old) * scale (:96); after the row every y_j of the row moves by
// lr * (steps_f - regImp * y_jf * n) and the loss takes regImp * y_jf^2 * n (:99-106).  loss *= 0.5.
// -------------------------------------------------------------------------------------
LRO_API double lro_svdpp_epoch(int32_t U, const int64_t* rowptr, const int32_t* col, const double* val, int32_t k,
                               double* P, double* Q, double* Y, double* bu, double* bi, double mu,
                               float lr_f, float regU_f, float regI_f, double regB, double regImp) {
    const double learnRate = (double)lr_f, regUser = (double)regU_f, regItem = (double)regI_f;
    double loss = 0.0;
    std::vector<double> fv((size_t)k), steps((size_t)k);
    for (int32_t u = 0; u < U; ++u) {
        const int64_t b = rowptr[u], e = rowptr[u + 1];
        const int64_t n = e - b;
        if (n == 0) continue;
        std::fill(steps.begin(), steps.end(), 0.0);
        std::fill(fv.begin(), fv.end(), 0.0);
        for (int64_t t = b; t < e; ++t)
            for (int f = 0; f < k; ++f) fv[(size_t)f] = Y[(int64_t)col[t] * k + f] + fv[(size_t)f];
        const double scale = pow((double)n, -0.5);
        for (int f = 0; f < k; ++f) fv[(size_t)f] = fv[(size_t)f] * scale;
        double* pu = P + (int64_t)u * k;
        for (int64_t t = b; t < e; ++t) {
            const int32_t i = col[t];
            double* qi = Q + (int64_t)i * k;
            double pred = bu[u] + bi[i] + mu;
            for (int f = 0; f < k; ++f) pred += (fv[(size_t)f] + pu[f]) * qi[f];
            const double error = val[t] - pred;
            loss += error * error;
            const double ub = bu[u];
            bu[u] += learnRate * (error - regB * ub);
            loss += regB * ub * ub;
            const double ib = bi[i];
            bi[i] += learnRate * (error - regB * ib);
            loss += regB * ib * ib;
            for (int f = 0; f < k; ++f) {
                const double uf = pu[f], itf = qi[f];
                pu[f] += learnRate * (error * itf - regUser * uf);
                qi[f] += learnRate * (error * (uf + fv[(size_t)f]) - regItem * itf);
                loss += regUser * uf * uf + regItem * itf * itf;
                steps[(size_t)f] += error * itf * scale;
            }
        }
        const int32_t size = (int32_t)n;
        for (int64_t t = b; t < e; ++t) {
            double* yj = Y + (int64_t)col[t] * k;
            for (int f = 0; f < k; ++f) {
                const double factor = yj[f];
                yj[f] += learnRate * (steps[(size_t)f] - regImp * factor * size);
                loss += regImp * factor * factor * size;
            }
        }
    }
    return 0.5 * loss;
}

// SVDPlusPlusRecommender.predict(u, i) (:138-150): fv = sum y_j / Math.sqrt(n) (division, not the pow of training), then :126-135
LRO_API void lro_svdpp_predict_pairs(int32_t k, const double* P, const double* Q, const double* Y, const double* bu, const double* bi,
                                     double mu, const int64_t* tr_rowptr, const int32_t* tr_col, const int32_t* users,
                                     const int32_t* items, int64_t n, double* out) {
    std::vector<double> fv((size_t)k);
    for (int64_t t = 0; t < n; ++t) {
        const int32_t u = users[t], i = items[t];
        std::fill(fv.begin(), fv.end(), 0.0);
        const int64_t b = tr_rowptr[u], e = tr_rowptr[u + 1];
        for (int64_t x = b; x < e; ++x)
            for (int f = 0; f < k; ++f) fv[(size_t)f] = Y[(int64_t)tr_col[x] * k + f] + fv[(size_t)f];
        const double scale = sqrt((double)(e - b));
        if (scale > 0) for (int f = 0; f < k; ++f) fv[(size_t)f] = fv[(size_t)f] / scale;
        double v = bu[u] + bi[i] + mu;
        for (int f = 0; f < k; ++f) v += (fv[(size_t)f] + P[(int64_t)u * k + f]) * Q[(int64_t)i * k + f];
        out[t] = v;
    }
}

// Maths.logistic: math/algorithm/Maths.java:127-129  (1 / (1 + Math.exp(-x)))
static inline double logistic(double x) { return 1.0 / (1.0 + exp(-x)); }

// One BPR epoch: recommender/cf/ranking/BPRRecommender.java:48-93.  Draws come from the global
// java.util.Random, interleaved with updates.  Membership test (fastutil IntOpenHashSet,
// :101-112) is restated as a binary search in the sorted row -- same truth value.
// When `trip` is non-null the (u,i,j) triples are taken from it instead of the RNG
// (3*n int32) -- NOT reference behaviour; used to feed the GPU kernel's own samples through
// the reference arithmetic.  When `trip_out` is non-null the drawn triples are recorded.
LRO_API double lro_bpr_epoch(int32_t U, int32_t I, const int64_t* rowptr, const int32_t* col, int32_t k,
                             double* P, double* Q, float lr_f, float regU_f, float regI_f,
                             int64_t n_samples, const int32_t* trip, int32_t* trip_out) {
    const double learnRate = (double)lr_f, regUser = (double)regU_f, regItem = (double)regI_f;
    double loss = 0.0;
    for (int64_t s = 0; s < n_samples; ++s) {
        int32_t u, pi, nj;
        if (trip) { u = trip[3 * s]; pi = trip[3 * s + 1]; nj = trip[3 * s + 2]; }
        else {
            for (;;) {
                u = lro_uniform_int(U);
                const int64_t b = rowptr[u], e = rowptr[u + 1];
                const int64_t len = e - b;
                if (len == 0 || len == I) continue;
                pi = col[b + lro_uniform_int((int32_t)len)];
                do { nj = lro_uniform_int(I); } while (std::binary_search(col + b, col + e, nj));
                break;
            }
        }
        if (trip_out) { trip_out[3 * s] = u; trip_out[3 * s + 1] = pi; trip_out[3 * s + 2] = nj; }
        double* pu = P + (int64_t)u * k;
        double* qi = Q + (int64_t)pi * k;
        double* qj = Q + (int64_t)nj * k;
        const double diff = dot_lr(pu, qi, k) - dot_lr(pu, qj, k);
        loss += -log(logistic(diff));
        const double deri = logistic(-diff);
        for (int f = 0; f < k; ++f) {
            const double uf = pu[f], pf = qi[f], nf = qj[f];
            pu[f] += learnRate * (deri * (pf - nf) - regUser * uf);
            qi[f] += learnRate * (deri * uf - regItem * pf);
            qj[f] += learnRate * (deri * (-uf) - regItem * nf);
            loss += regUser * uf * uf + regItem * pf * pf + regItem * nf * nf;
        }
    }
    return loss;  // no *0.5 for BPR
}

// -------------------------------------------------------------------------------------
// RankSGD: recommender/cf/ranking/RankSGDRecommender.java:42-108.
// setup (:42-58): itemProbs = the items with at least one train rating, prob = users(j) / numRates, put into a
// java.util.HashMap<Integer, Double> in ascending item order and sorted ASCENDING by prob with the stable
// Collections.sort (util/Lists.java:266-308) -- so ties keep the HashMap's iteration order (ascending bucket
// (h ^ h>>>16) & (cap-1); the JDK helpers further down restate it; forward-declared here).
// -------------------------------------------------------------------------------------
static inline uint32_t jhash_bucket(int32_t key, uint32_t cap);
static inline uint32_t jhashset_capacity(int64_t n);

// out_items / out_probs: capacity I; returns the list length m
LRO_API int32_t lro_ranksgd_item_probs(int32_t U, int32_t I, const int64_t* rowptr, const int32_t* col,
                                       int32_t* out_items, double* out_probs) {
    const int64_t nnz = rowptr[U];
    std::vector<int64_t> users((size_t)I, 0);
    for (int64_t e = 0; e < nnz; ++e) users[(size_t)col[e]]++;
    struct Ent { int32_t item; double prob; uint64_t pos; };
    std::vector<Ent> ents;
    int64_t m = 0;
    for (int32_t j = 0; j < I; ++j) if (users[(size_t)j] > 0) ++m;
    const uint32_t cap = jhashset_capacity(m);
    uint64_t seq = 0;
    for (int32_t j = 0; j < I; ++j) {
        const double prob = ((double)users[(size_t)j] + 0.0) / (double)nnz;      // (users + 0.0) / numRates
        if (prob > 0) ents.push_back({j, prob, ((uint64_t)jhash_bucket(j, cap) << 32) | (seq++)});
    }
    std::sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b) { return a.pos < b.pos; });       // entrySet() order
    std::stable_sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b) { return a.prob < b.prob; });
    for (size_t t = 0; t < ents.size(); ++t) { out_items[t] = ents[t].item; out_probs[t] = ents[t].prob; }
    return (int32_t)ents.size();
}

// One RankSGD epoch (:62-108): every train entry in CSR order; the negative is drawn by walking the cumulative sum of
// the ascending list until sum >= rand (Randoms.random() = nextDouble), redrawn while the user has rated it; if the
// walk ends without a hit (sum of probs < rand by rounding) negItemIdx keeps its previous value, -1 at the first try
// of an entry -- the reference would then index itemFactors with -1 and throw; restated as "return NaN".
// error = (pos - neg) - (r - 0); loss += error^2; NO regularisation; old user factor feeds both item updates; loss *= 0.5.
// trip (3*n int32 (u, i, j)) replaces the CSR walk and the RNG -- NOT reference behaviour, used to run the GPU kernel's own
// samples through this arithmetic; the rating of (u, i) is looked up in the CSR.  trip_out records what was drawn.
LRO_API double lro_ranksgd_epoch(int32_t U, int32_t I, const int64_t* rowptr, const int32_t* col, const double* val, int32_t k,
                                 double* P, double* Q, float lr_f, int64_t n, const int32_t* trip, int32_t* trip_out) {
    const double learnRate = (double)lr_f;
    std::vector<int32_t> items((size_t)I);
    std::vector<double> probs((size_t)I);
    int32_t m = 0;
    if (!trip) m = lro_ranksgd_item_probs(U, I, rowptr, col, items.data(), probs.data());
    double loss = 0.0;
    int32_t u = 0;
    const int64_t total = trip ? n : rowptr[U];
    for (int64_t t = 0; t < total; ++t) {
        int32_t pi, nj;
        double r;
        if (trip) {
            u = trip[3 * t]; pi = trip[3 * t + 1]; nj = trip[3 * t + 2];
            const int32_t* b = col + rowptr[u];
            const int32_t* e = col + rowptr[u + 1];
            const int32_t* f = std::lower_bound(b, e, pi);
            if (f == e || *f != pi) return std::nan("");
            r = val[f - col];
        } else {
            while (rowptr[u + 1] <= t) ++u;
            pi = col[t]; r = val[t];
            nj = -1;
            for (;;) {
                double sum = 0;
                const double rand = lro_uniform();
                for (int32_t c = 0; c < m; ++c) {
                    sum += probs[(size_t)c];
                    if (sum >= rand) { nj = items[(size_t)c]; break; }
                }
                if (nj < 0 || !std::binary_search(col + rowptr[u], col + rowptr[u + 1], nj)) break;
            }
            if (nj < 0) return std::nan("");
        }
        if (trip_out) { trip_out[3 * t] = u; trip_out[3 * t + 1] = pi; trip_out[3 * t + 2] = nj; }
        double* pu = P + (int64_t)u * k;
        double* qi = Q + (int64_t)pi * k;
        double* qj = Q + (int64_t)nj * k;
        const double error = (dot_lr(pu, qi, k) - dot_lr(pu, qj, k)) - (r - 0.0);
        loss += error * error;
        const double sgd = learnRate * error;
        for (int f = 0; f < k; ++f) {
            const double uf = pu[f], pf = qi[f], nf = qj[f];
            pu[f] += -sgd * (pf - nf);
            qi[f] += -sgd * uf;
            qj[f] += sgd * uf;
        }
    }
    return 0.5 * loss;
}

// -------------------------------------------------------------------------------------
// AoBPR (SURVEY.md 8f, row N3 -- groundwork: oracle only): recommender/cf/ranking/AoBPRRecommender.java:51-216.
// BPR whose negative item is drawn by RANK: setup (:51-73) lambdaItem = (int)(rec.item.distribution.parameter * numItems),
// loopNumber = (int)(numItems * ln numItems), RankingPro[i] = exp(-((i + 1) / lambdaItem)) with INTEGER division (a step
// function -- kept), normalised.  Every loopNumber samples (counter carried over the epochs, :88-92) each factor's items are
// re-sorted by value, descending and stable (:186-203), and var[f] is the population variance of that column.  A sample
// (:95-130): train entry uniform(numRates) -> (u, i); rank r = discrete(RankingPro); factor f = discrete(|p_uf| var[f] / sum);
// j = factorRanking[f][r] if p_uf > 0 else factorRanking[f][numItems - r - 1]; redraw (r and f) while u has rated j.
// Randoms.discrete (math/algorithm/Randoms.java:468-489): one uniform per attempt, first i with cumulative sum > r.
// The update is BPR's (:133-148), no 0.5 on the loss.  losses_out[iter]; trip_out (3 * nnz, optional) records iteration 1.
// returns 0, or -1 when discrete() would throw (probabilities not summing to one, e.g. all-zero user factors).
// -------------------------------------------------------------------------------------
static int32_t lro_discrete(const double* a, int32_t n) {
    double sum = 0.0;
    for (int32_t i = 0; i < n; ++i) { if (a[i] < 0.0) return -1; sum = sum + a[i]; }
    if (!(sum <= 1.0 + 1E-6 && sum >= 1.0 - 1E-6)) return -1;
    for (;;) {
        const double r = lro_uniform();
        sum = 0.0;
        for (int32_t i = 0; i < n; ++i) { sum = sum + a[i]; if (sum > r) return i; }
    }
}
LRO_API int32_t lro_aobpr_train(int32_t U, int32_t I, const int64_t* rowptr, const int32_t* col, int32_t k, double* P, double* Q,
                                float lr_f, float regU_f, float regI_f, float dist_param, int32_t num_iter, double* losses_out,
                                int32_t* trip_out) {
    const double learnRate = (double)lr_f, regUser = (double)regU_f, regItem = (double)regI_f;
    const int64_t nnz = rowptr[U];
    const int32_t lambdaItem = (int32_t)(dist_param * (float)I);
    const int32_t loopNumber = (int32_t)((double)I * log((double)I));
    if (lambdaItem == 0 || loopNumber == 0) return -1;                       // ArithmeticException in the reference
    std::vector<double> rankingPro((size_t)I), var((size_t)k), pfc((size_t)k), values((size_t)I);
    {
        double sum = 0;
        for (int32_t i = 0; i < I; ++i) { rankingPro[(size_t)i] = exp((double)(-((i + 1) / lambdaItem))); sum += rankingPro[(size_t)i]; }
        for (int32_t i = 0; i < I; ++i) rankingPro[(size_t)i] /= sum;
    }
    std::vector<int32_t> ranking((size_t)k * I), user_of((size_t)nnz);
    for (int32_t u = 0; u < U; ++u) for (int64_t e = rowptr[u]; e < rowptr[u + 1]; ++e) user_of[(size_t)e] = u;
    int64_t countIter = 0;
    for (int32_t iter = 1; iter <= num_iter; ++iter) {
        double loss = 0.0;
        for (int64_t s = 0; s < nnz; ++s) {
            if (countIter % loopNumber == 0) {
                for (int32_t f = 0; f < k; ++f) {
                    int32_t* rk = ranking.data() + (size_t)f * I;
                    for (int32_t i = 0; i < I; ++i) rk[i] = i;
                    std::stable_sort(rk, rk + I, [&](int32_t a, int32_t b) { return Q[(int64_t)a * k + f] > Q[(int64_t)b * k + f]; });
                    double m = 0.0;
                    for (int32_t i = 0; i < I; ++i) { values[(size_t)i] = Q[(int64_t)rk[i] * k + f]; m += values[(size_t)i]; }
                    m = m / I;
                    double v = 0.0;
                    for (int32_t i = 0; i < I; ++i) v += (values[(size_t)i] - m) * (values[(size_t)i] - m);
                    var[(size_t)f] = v / (I - 0);
                }
                countIter = 0;
            }
            countIter++;
            int32_t u, pi, nj;
            for (;;) {
                const int32_t dataIdx = lro_uniform_int((int32_t)nnz);
                u = user_of[(size_t)dataIdx];
                const int64_t b = rowptr[u], e = rowptr[u + 1];
                if (e - b == 0 || e - b == I) continue;
                pi = col[dataIdx];
                do {
                    int32_t r;
                    do { r = lro_discrete(rankingPro.data(), I); if (r < 0) return -1; } while (r > I);
                    double sumfc = 0;
                    for (int32_t f = 0; f < k; ++f) {
                        const double t = fabs(P[(int64_t)u * k + f]);
                        sumfc += t * var[(size_t)f];
                        pfc[(size_t)f] = t * var[(size_t)f];
                    }
                    for (int32_t f = 0; f < k; ++f) pfc[(size_t)f] /= sumfc;
                    const int32_t f = lro_discrete(pfc.data(), k);
                    if (f < 0) return -1;
                    nj = P[(int64_t)u * k + f] > 0 ? ranking[(size_t)f * I + r] : ranking[(size_t)f * I + (I - r - 1)];
                } while (std::binary_search(col + b, col + e, nj));
                break;
            }
            if (trip_out && iter == 1) { trip_out[3 * s] = u; trip_out[3 * s + 1] = pi; trip_out[3 * s + 2] = nj; }
            double* pu = P + (int64_t)u * k;
            double* qi = Q + (int64_t)pi * k;
            double* qj = Q + (int64_t)nj * k;
            const double diff = dot_lr(pu, qi, k) - dot_lr(pu, qj, k);
            loss += -log(logistic(diff));
            const double deri = logistic(-diff);
            for (int f = 0; f < k; ++f) {
                const double uf = pu[f], pf = qi[f], nf = qj[f];
                pu[f] += learnRate * (deri * (pf - nf) - regUser * uf);
                qi[f] += learnRate * (deri * uf - regItem * pf);
                qj[f] += learnRate * (deri * (-uf) - regItem * nf);
                loss += regUser * uf * uf + regItem * pf * pf + regItem * nf * nf;
            }
        }
        if (losses_out) losses_out[iter - 1] = loss;
    }
    return 0;
}

// -------------------------------------------------------------------------------------
// GBPR (SURVEY.md 8f, row N3 -- groundwork: oracle only): recommender/cf/ranking/GBPRRecommender.java:56-196.
// Group preference: a sample is (u, i, G, j) -- u uniform over users with ratings, i uniform in u's row, G a java.util.HashSet
// of gLen users who rated i (all of them when at most gLen did; otherwise u plus uniform draws from the column until the set
// is full, :95-105), j uniform over the items u has not rated.  Prediction of (u, i, G) = rho * (mean_g p_g.q_i + b_i) +
// (1 - rho) * (b_i + p_u.q_i) (:185-192); loss and derivative as BPR.  The item biases move immediately (:121-126); the
// FACTOR updates are accumulated in temporaries and applied at the END of the epoch (:68-69,160-161) -- within an epoch every
// sample sees the epoch-start factors.  The group is walked in HashSet iteration order (sumGroup and the loss are sums in
// that order).  rho is a float (default 1.5f): (1 - rho) is float arithmetic.  No 0.5 on the loss.
// group_out (optional, n * gLen int32, -1 padded) / trip_out (3 * n) record what was drawn.
// -------------------------------------------------------------------------------------
LRO_API double lro_gbpr_epoch(int32_t U, int32_t I, const int64_t* rowptr, const int32_t* col, int32_t k, double* P, double* Q, double* bi,
                              float lr_f, float regU_f, float regI_f, double regB, float rho, int32_t g_len,
                              int32_t* trip_out, int32_t* group_out) {
    const double learnRate = (double)lr_f, regUser = (double)regU_f, regItem = (double)regI_f;
    const int64_t nnz = rowptr[U];
    std::vector<int64_t> colptr, csc;
    lro_csc_order(U, I, rowptr, col, colptr, csc);
    std::vector<int32_t> user_of((size_t)nnz);
    for (int32_t u = 0; u < U; ++u) for (int64_t e = rowptr[u]; e < rowptr[u + 1]; ++e) user_of[(size_t)e] = u;
    std::vector<double> tP((size_t)U * k, 0.0), tQ((size_t)I * k, 0.0), sumGroup((size_t)k);
    std::vector<int32_t> group, order;
    const double one_minus_rho = (double)(1 - rho);                        // float subtraction, then promotion
    double loss = 0.0;
    for (int64_t s = 0; s < nnz; ++s) {
        int32_t u;
        int64_t b, e;
        do { u = lro_uniform_int(U); b = rowptr[u]; e = rowptr[u + 1]; } while (e - b == 0);
        const int32_t pi = col[b + lro_uniform_int((int32_t)(e - b))];
        // users group: insertion order first, then HashSet iteration order
        const int64_t cb = colptr[(size_t)pi], ce = colptr[(size_t)pi + 1];
        group.clear();
        if (ce - cb <= g_len) {
            for (int64_t x = cb; x < ce; ++x) group.push_back(user_of[(size_t)csc[(size_t)x]]);
        } else {
            group.push_back(u);
            while ((int32_t)group.size() < g_len) {
                const int32_t t = user_of[(size_t)csc[(size_t)(cb + lro_uniform_int((int32_t)(ce - cb)))]];
                if (std::find(group.begin(), group.end(), t) == group.end()) group.push_back(t);
            }
        }
        {
            const uint32_t cap = jhashset_capacity((int64_t)group.size());
            std::vector<std::pair<uint64_t, int32_t>> ord;
            for (size_t x = 0; x < group.size(); ++x) ord.push_back({((uint64_t)jhash_bucket(group[x], cap) << 32) | (uint32_t)x, group[x]});
            std::sort(ord.begin(), ord.end());
            order.clear();
            for (auto& o : ord) order.push_back(o.second);
        }
        const double* qi = Q + (int64_t)pi * k;
        const double predictRating = bi[pi] + dot_lr(P + (int64_t)u * k, qi, k);
        double sum = 0;
        for (int32_t g : order) sum += dot_lr(P + (int64_t)g * k, qi, k);
        const double groupRating = sum / (double)order.size() + bi[pi];
        const double posPredict = rho * groupRating + one_minus_rho * predictRating;
        int32_t nj;
        do { nj = lro_uniform_int(I); } while (std::binary_search(col + b, col + e, nj));
        const double* qj = Q + (int64_t)nj * k;
        const double negPredict = bi[nj] + dot_lr(P + (int64_t)u * k, qj, k);
        if (trip_out) { trip_out[3 * s] = u; trip_out[3 * s + 1] = pi; trip_out[3 * s + 2] = nj; }
        if (group_out) for (int32_t x = 0; x < g_len; ++x) group_out[s * g_len + x] = x < (int32_t)order.size() ? order[(size_t)x] : -1;
        const double diff = posPredict - negPredict;
        loss += -log(logistic(diff));
        const double deri = logistic(-diff);
        const double pb = bi[pi];
        bi[pi] += learnRate * (deri - regB * pb);
        loss += regB * pb * pb;
        const double nb = bi[nj];
        bi[nj] += learnRate * (-deri - regB * nb);
        loss += regB * nb * nb;
        const double avgW = 1.0 / (double)order.size();
        std::fill(sumGroup.begin(), sumGroup.end(), 0.0);
        for (int32_t g : order) {
            const double delta = g == u ? 1 : 0;
            for (int f = 0; f < k; ++f) {
                const double gf = P[(int64_t)g * k + f], pf = qi[f], nf = qj[f];
                const double deltaGroup = rho * avgW * pf + one_minus_rho * delta * pf - delta * nf;
                tP[(size_t)g * k + f] += learnRate * (deri * deltaGroup - regUser * gf);
                loss += regUser * gf * gf;
                sumGroup[(size_t)f] += gf;
            }
        }
        for (int f = 0; f < k; ++f) {
            const double uf = P[(int64_t)u * k + f], pf = qi[f], nf = qj[f];
            const double posDelta = rho * avgW * sumGroup[(size_t)f] + one_minus_rho * uf;
            tQ[(size_t)pi * k + f] += learnRate * (deri * posDelta - regItem * pf);
            loss += regItem * pf * pf;
            loss += regItem * nf * nf;
            const double negDelta = -uf;
            tQ[(size_t)nj * k + f] += learnRate * (deri * negDelta - regItem * nf);
        }
    }
    for (size_t t = 0; t < tP.size(); ++t) P[t] = P[t] + tP[t];
    for (size_t t = 0; t < tQ.size(); ++t) Q[t] = Q[t] + tQ[t];
    return loss;
}

// -------------------------------------------------------------------------------------
// ALS siblings: WRMF and eALS.  Both are deterministic (no RNG inside trainModel; every row's update
// reads only the OTHER side's matrix, so the reference's parallelStream order does not matter), which makes them the two models
// of the path whose device results can be held to BIT equality with this restatement.
//
// math/structure/DenseMatrix.java:362-437 -- inverse(): Gauss-Jordan with partial pivoting on a clone, the inverse built in a
// second matrix.  Quirks kept: the pivot search takes the FIRST strictly-largest |a_jr| (NaN never wins); "no pivot" returns the
// half-built inverse as it stands; a row swap exchanges only columns >= r of the clone; the pivot row is divided (not multiplied
// by a reciprocal); the elimination factor is read before the row is touched.
static void lro_dense_inverse(const double* a, int32_t n, double* inv) {
    for (int32_t r = 0; r < n; ++r) for (int32_t c = 0; c < n; ++c) inv[(size_t)r * n + c] = r == c ? 1.0 : 0.0;
    if (n == 1) { inv[0] = 1.0 / a[0]; return; }
    std::vector<double> m(a, a + (size_t)n * n);
    for (int32_t r = 0; r < n; ++r) {
        double mag = 0.0;
        int32_t pivot = -1;
        for (int32_t j = r; j < n; ++j) {
            const double mag2 = fabs(m[(size_t)j * n + r]);
            if (mag2 > mag) { mag = mag2; pivot = j; }
        }
        if (pivot == -1 || mag == 0) return;
        if (pivot != r) {
            for (int32_t c = r; c < n; ++c) std::swap(m[(size_t)r * n + c], m[(size_t)pivot * n + c]);
            for (int32_t c = 0; c < n; ++c) std::swap(inv[(size_t)r * n + c], inv[(size_t)pivot * n + c]);
        }
        mag = m[(size_t)r * n + r];
        for (int32_t c = r; c < n; ++c) m[(size_t)r * n + c] = m[(size_t)r * n + c] / mag;
        for (int32_t c = 0; c < n; ++c) inv[(size_t)r * n + c] = inv[(size_t)r * n + c] / mag;
        for (int32_t r2 = 0; r2 < n; ++r2) {
            if (r == r2) continue;
            const double mag2 = m[(size_t)r2 * n + r];
            for (int32_t c = r; c < n; ++c) m[(size_t)r2 * n + c] = m[(size_t)r2 * n + c] - mag2 * m[(size_t)r * n + c];
            for (int32_t c = 0; c < n; ++c) inv[(size_t)r2 * n + c] = inv[(size_t)r2 * n + c] - mag2 * inv[(size_t)r * n + c];
        }
    }
}
LRO_API void lro_dense_inverse_export(const double* a, int32_t n, double* inv) { lro_dense_inverse(a, n, inv); }

// M^T M through DenseMatrix.transpose().times(M) (DenseMatrix.java:229-249,276-286; DenseVector.java:104-111): entry (r, c) is
// the dot of row r of M^T with column c of M, summed over the rows of M in order, starting from 0.0
static void lro_gram(const double* M, int64_t n, int32_t k, double* out) {
    for (int32_t r = 0; r < k; ++r)
        for (int32_t c = 0; c < k; ++c) {
            double v = 0.0;
            for (int64_t i = 0; i < n; ++i) v += M[i * k + c] * M[i * k + r];
            out[(size_t)r * k + c] = v;
        }
}

// WRMFRecommender.weight (recommender/cf/ranking/WRMFRecommender.java:58-61): log(1 + 10^coef * value); coef is a float.
// The reference calls Math.log / Math.pow (intrinsics, allowed to differ from StrictMath by an ulp); the device never computes it --
// the weighting stays in the Java shim's setup(), like weightMatrix() (:63-72).
LRO_API double lro_wrmf_weight(double value, float coef) { return fdlibm_log(1.0 + pow(10.0, (double)coef) * value); }

// one half-iteration of WRMFRecommender.trainModel (:93-126 users, :129-163 items): for every row of `ptr/idx/w` (the weighted train
// matrix by rows, or by columns for the item step) solve (F^T F + reg [EVERY entry, :110] + sum_e w_e f_e f_e^T) x = sum_e (w_e + 1) f_e
static void lro_wrmf_side(int32_t n_rows, const int64_t* ptr, const int32_t* idx, const double* w, int32_t k, const double* F, int64_t n_other,
                          double* X, double reg) {
    std::vector<double> G((size_t)k * k), A((size_t)k * k), W((size_t)k * k), b((size_t)k);
    lro_gram(F, n_other, k, G.data());
    for (int32_t r = 0; r < n_rows; ++r) {
        std::fill(b.begin(), b.end(), 0.0);
        for (int64_t e = ptr[r]; e < ptr[r + 1]; ++e) {
            const double weight = w[e] + 1.0;
            const double* f = F + (int64_t)idx[e] * k;
            for (int32_t a = 0; a < k; ++a) b[(size_t)a] += f[a] * weight;
        }
        for (size_t t = 0; t < A.size(); ++t) A[t] = G[t] + reg;
        for (int64_t e = ptr[r]; e < ptr[r + 1]; ++e) {
            const double weight = w[e];
            const double* f = F + (int64_t)idx[e] * k;
            for (int32_t a = 0; a < k; ++a) {
                const double temp = f[a] * weight;
                for (int32_t c = 0; c < k; ++c) A[(size_t)a * k + c] += temp * f[c];
            }
        }
        lro_dense_inverse(A.data(), k, W.data());
        for (int32_t a = 0; a < k; ++a) {                     // DenseMatrix.times(Vector): row(a).dot(b)
            double v = 0.0;
            for (int32_t c = 0; c < k; ++c) v += b[(size_t)c] * W[(size_t)a * k + c];
            X[(int64_t)r * k + a] = v;
        }
    }
}

// one iteration of WRMFRecommender.trainModel; val = the WEIGHTED train values (weightMatrix has run); the model has no loss
LRO_API void lro_wrmf_epoch(int32_t U, int32_t I, const int64_t* rowptr, const int32_t* col, const double* val, int32_t k,
                            double* P, double* Q, float regU_f, float regI_f) {
    std::vector<int64_t> colptr, csc;
    lro_csc_order(U, I, rowptr, col, colptr, csc);
    const int64_t nnz = rowptr[U];
    std::vector<int32_t> cusers((size_t)nnz);
    std::vector<double> cval((size_t)nnz);
    {
        std::vector<int32_t> user_of((size_t)nnz);
        for (int32_t u = 0; u < U; ++u) for (int64_t e = rowptr[u]; e < rowptr[u + 1]; ++e) user_of[(size_t)e] = u;
        for (int64_t x = 0; x < nnz; ++x) { cusers[(size_t)x] = user_of[(size_t)csc[(size_t)x]]; cval[(size_t)x] = val[csc[(size_t)x]]; }
    }
    lro_wrmf_side(U, rowptr, col, val, k, Q, I, P, (double)regU_f);
    lro_wrmf_side(I, colptr.data(), cusers.data(), cval.data(), k, P, U, Q, (double)regI_f);
}

// eALS (recommender/cf/ranking/EALSRecommender.java:53-112 setup, :114-214 trainModel).
// confidences (:65-83): judge 0 or 2: c_i = overall * pop_i^ratio / sum_j pop_j^ratio with pop_i = |column i| / numRates; else 1.
// weight(value) (:85-95): judge 1 or 2: 1 + coef * value (coef float); else 1.
LRO_API void lro_eals_confidences(int32_t U, int32_t I, const int64_t* rowptr, const int32_t* col, float ratio, float overall, int32_t judge,
                                  double* conf) {
    if (judge == 0 || judge == 2) {
        std::vector<int64_t> cnt((size_t)I, 0);
        const int64_t nnz = rowptr[U];
        for (int64_t e = 0; e < nnz; ++e) cnt[(size_t)col[e]]++;
        double sumPopularity = 0.0;
        for (int32_t i = 0; i < I; ++i) {
            const double alphaPopularity = pow((double)cnt[(size_t)i] * 1.0 / (double)nnz, (double)ratio);
            conf[i] = (double)overall * alphaPopularity;
            sumPopularity += alphaPopularity;
        }
        for (int32_t i = 0; i < I; ++i) conf[i] = conf[i] / sumPopularity;
    } else
        for (int32_t i = 0; i < I; ++i) conf[i] = 1;
}
LRO_API double lro_eals_weight(double value, float coef, int32_t judge) { return judge == 1 || judge == 2 ? 1.0 + (double)coef * value : 1.0; }

// one iteration of EALSRecommender.trainModel (:127-211).  The caller zeroes P before the first one (:125 replaces userFactors by a
// fresh zero matrix).  val = weighted train values; element-wise coordinate updates, every sum in the reference's order.
LRO_API void lro_eals_epoch(int32_t U, int32_t I, const int64_t* rowptr, const int32_t* col, const double* val, int32_t k,
                            double* P, double* Q, const double* conf, float regU_f, float regI_f) {
    const double regUser = (double)regU_f, regItem = (double)regI_f;
    const int64_t nnz = rowptr[U];
    std::vector<int64_t> colptr, csc;
    lro_csc_order(U, I, rowptr, col, colptr, csc);
    std::vector<int32_t> user_of((size_t)nnz);
    for (int32_t u = 0; u < U; ++u) for (int64_t e = rowptr[u]; e < rowptr[u + 1]; ++e) user_of[(size_t)e] = u;
    std::vector<double> Sq((size_t)k * k), Sp((size_t)k * k), itemsPredictions((size_t)I, 0.0), usersPredictions((size_t)U, 0.0);
    for (int32_t f1 = 0; f1 < k; ++f1)
        for (int32_t f2 = 0; f2 <= f1; ++f2) {
            double value = 0;
            for (int32_t i = 0; i < I; ++i) value += conf[i] * Q[(int64_t)i * k + f1] * Q[(int64_t)i * k + f2];
            Sq[(size_t)f1 * k + f2] = value;
            Sq[(size_t)f2 * k + f1] = value;
        }
    for (int32_t u = 0; u < U; ++u) {
        const int64_t b = rowptr[u], e = rowptr[u + 1];
        double* pu = P + (int64_t)u * k;
        for (int64_t x = b; x < e; ++x) {
            const double* qi = Q + (int64_t)col[x] * k;
            double d = 0.0;
            for (int32_t f = 0; f < k; ++f) d += qi[f] * pu[f];
            itemsPredictions[(size_t)col[x]] = d;
        }
        for (int32_t f = 0; f < k; ++f) {
            double numer = 0, denom = regUser + Sq[(size_t)f * k + f];
            for (int32_t f2 = 0; f2 < k; ++f2)
                if (f != f2) numer -= pu[f2] * Sq[(size_t)f * k + f2];
            for (int64_t x = b; x < e; ++x) {
                const int32_t i = col[x];
                const double weight = val[x], qf = Q[(int64_t)i * k + f];
                itemsPredictions[(size_t)i] -= pu[f] * qf;
                numer += (weight - (weight - conf[i]) * itemsPredictions[(size_t)i]) * qf;
                denom += (weight - conf[i]) * qf * qf;
            }
            pu[f] = numer / denom;
            for (int64_t x = b; x < e; ++x) {
                const int32_t i = col[x];
                itemsPredictions[(size_t)i] += pu[f] * Q[(int64_t)i * k + f];
            }
        }
    }
    lro_gram(P, U, k, Sp.data());
    for (int32_t i = 0; i < I; ++i) {
        const int64_t b = colptr[(size_t)i], e = colptr[(size_t)i + 1];
        double* qi = Q + (int64_t)i * k;
        for (int64_t x = b; x < e; ++x) {
            const int32_t u = user_of[(size_t)csc[(size_t)x]];
            const double* pu = P + (int64_t)u * k;
            double d = 0.0;
            for (int32_t f = 0; f < k; ++f) d += qi[f] * pu[f];
            usersPredictions[(size_t)u] = d;
        }
        for (int32_t f = 0; f < k; ++f) {
            double numer = 0, denom = conf[i] * Sp[(size_t)f * k + f] + regItem;
            for (int32_t f2 = 0; f2 < k; ++f2)
                if (f != f2) numer -= qi[f2] * Sp[(size_t)f2 * k + f];
            numer *= conf[i];
            for (int64_t x = b; x < e; ++x) {
                const int32_t u = user_of[(size_t)csc[(size_t)x]];
                const double weight = val[csc[(size_t)x]], pf = P[(int64_t)u * k + f];
                usersPredictions[(size_t)u] -= pf * qi[f];
                numer += (weight - (weight - conf[i]) * usersPredictions[(size_t)u]) * pf;
                denom += (weight - conf[i]) * pf * pf;
            }
            qi[f] = numer / denom;
            for (int64_t x = b; x < e; ++x) {
                const int32_t u = user_of[(size_t)csc[(size_t)x]];
                usersPredictions[(size_t)u] += P[(int64_t)u * k + f] * qi[f];
            }
        }
    }
}

// AbstractRecommender.isConverged: recommender/AbstractRecommender.java:249-267.
// returns 1 converged, 0 not, -1 = would throw LibrecException (NaN / Inf loss)
LRO_API int32_t lro_is_converged(double last_loss, double loss, float* delta_out) {
    float delta = (float)(last_loss - loss);
    if (delta_out) *delta_out = delta;
    if (std::isnan(loss) || std::isinf(loss)) return -1;
    return fabs((double)delta) < 1e-5 ? 1 : 0;   // Math.abs(float) widened against the double literal
}

// MatrixFactorizationRecommender.updateLRate: recommender/MatrixFactorizationRecommender.java:121-139
// (float arithmetic on learnRate).  Returns the new learn rate; *last_loss is updated.
LRO_API float lro_update_lrate(float learnRate, float maxLearnRate, int32_t iter, int32_t bold_driver, float decay,
                               double loss, double* last_loss) {
    if (learnRate < 0.0) { *last_loss = loss; return learnRate; }
    if (bold_driver && iter > 1) {
        learnRate = fabs(*last_loss) > fabs(loss) ? learnRate * 1.05f : learnRate * 0.5f;
    } else if (decay > 0 && decay < 1) {
        learnRate *= decay;
    }
    if (maxLearnRate > 0 && learnRate > maxLearnRate) learnRate = maxLearnRate;
    *last_loss = loss;
    return learnRate;
}

// Full trainModel loop (BiasedMF :67-107, PMF, BPR :45-99).  losses_out[iter-1] = loss of iter.
// returns iterations executed, or -iter when isConverged would throw at that iteration.
LRO_API int32_t lro_train(int32_t model, int32_t U, int32_t I, const int64_t* rowptr, const int32_t* col,
                          const double* val, int32_t k, double* P, double* Q, double* bu, double* bi, double mu,
                          float learnRate, float maxLearnRate, float regU, float regI, double regB,
                          int32_t num_iter, int32_t early_stop, int32_t bold_driver, float decay,
                          double* losses_out) {
    double last_loss = 0.0;   // AbstractRecommender.lastLoss initial value (:76)
    int32_t done = 0;
    for (int32_t iter = 1; iter <= num_iter; ++iter) {
        double loss;
        if (model == LRO_BIASEDMF) loss = lro_biasedmf_epoch(U, rowptr, col, val, k, P, Q, bu, bi, mu, learnRate, regU, regI, regB, nullptr, nullptr);
        else if (model == LRO_PMF) loss = lro_pmf_epoch(U, rowptr, col, val, k, P, Q, learnRate, regU, regI, nullptr, nullptr);
        else if (model == LRO_RANKSGD) loss = lro_ranksgd_epoch(U, I, rowptr, col, val, k, P, Q, learnRate, 0, nullptr, nullptr);
        else loss = lro_bpr_epoch(U, I, rowptr, col, k, P, Q, learnRate, regU, regI, rowptr[U], nullptr, nullptr);
        if (losses_out) losses_out[iter - 1] = loss;
        done = iter;
        int32_t c = lro_is_converged(last_loss, loss, nullptr);
        if (c < 0) return -iter;
        if (c == 1 && early_stop) break;
        learnRate = lro_update_lrate(learnRate, maxLearnRate, iter, bold_driver, decay, loss, &last_loss);
    }
    return done;
}

// -------------------------------------------------------------------------------------
// Rating prediction + RMSE / MAE:
// recommender/MatrixRecommender.java:211-248,272-284 (clamp to [minRate,maxRate]; NaN -> globalMean),
// eval/rating/RMSEEvaluator.java:33-69 (Math.pow(d,2) == d*d exactly), eval/rating/MAEEvaluator.java:34-70,
// both accumulated in test-CSR order.  pred_out (optional) receives the bounded predictions.
// -------------------------------------------------------------------------------------
LRO_API void lro_eval_rating(int32_t model, int32_t U, const int64_t* t_rowptr, const int32_t* t_col, const double* t_val,
                             int32_t k, const double* P, const double* Q, const double* bu, const double* bi, double mu,
                             double min_rate, double max_rate, double* rmse, double* mae, double* pred_out) {
    double se = 0.0, ae = 0.0;
    int64_t n = 0;
    for (int32_t u = 0; u < U; ++u)
        for (int64_t e = t_rowptr[u]; e < t_rowptr[u + 1]; ++e) {
            double p = predict_raw(model, k, P, Q, bu, bi, mu, u, t_col[e]);
            if (p > max_rate) p = max_rate; else if (p < min_rate) p = min_rate;
            if (std::isnan(p)) p = mu;
            if (pred_out) pred_out[e] = p;
            const double d = t_val[e] - p;
            se += d * d; ae += fabs(d); ++n;
        }
    *rmse = n > 0 ? sqrt(se / (double)n) : 0.0;
    *mae = n > 0 ? ae / (double)n : 0.0;
}

// -------------------------------------------------------------------------------------
// Ranking evaluators over top-N lists:
//   eval/ranking/AUCEvaluator.java:45-106, AveragePrecisionEvaluator.java:43-72,
//   NormalizedDCGEvaluator.java:44-103 (+ getValueByKey :105-113), PrecisionEvaluator.java:24-47,
//   RecallEvaluator.java:44-66, ReciprocalRankEvaluator.java:43-64.
// Ground truth of a user = its test row in CSR order (eval/EvalContext.java:75-88); a user counts only if that
// row is not empty.  numDropped[u] = numItems - |train row u| (recommender/MatrixRecommender.java:110-113).
// JDK behaviour the reference leans on (not vendored): the AUC pair count iterates a java.util.HashSet<Integer>
// built by adding the test items in ascending order -- iteration order = ascending bucket
// ((h ^ h>>>16) & (cap-1), cap = 16 doubled while size > 0.75 cap), insertion order inside a bucket.
// Java quirks kept: Precision divides by topN, not by the list length; AP divides by min(|test|, topK);
// "NDCG" takes its ideal DCG only from the ground-truth entries that were hit.
// out[8] = {AUC, AP, NDCG, Precision, Recall, RR, Novelty, Entropy}; the first six are means over the users that count.
// Novelty (NoveltyEvaluator.java:62-84): self-information of the recommended items under the purchase probability
// count_i / numUsers, count_i = train + test column counts (MatrixRecommender.java:118-122), summed over ALL users'
// lists, / (numUsers ln 2).  Entropy (EntropyEvaluator.java:60-90): entropy in bits of the items' frequency in the lists.
// -------------------------------------------------------------------------------------
static inline uint32_t jhash_bucket(int32_t key, uint32_t cap) {
    const uint32_t h = (uint32_t)key;
    return (h ^ (h >> 16)) & (cap - 1);
}
static inline uint32_t jhashset_capacity(int64_t n) {          // HashMap.putVal / resize with the default load factor
    uint32_t cap = 16;
    while ((double)n > 0.75 * (double)cap) cap <<= 1;
    return cap;
}
LRO_API void lro_eval_ranking(int32_t U, int32_t topn, const int32_t* rec_items, const int32_t* rec_counts,
                              const int64_t* t_rowptr, const int32_t* t_col, const double* t_val,
                              const int32_t* num_dropped, const int32_t* item_purchase_num, int32_t I, double* out) {
    double auc = 0, ap = 0, ndcg = 0, prec = 0, rec = 0, rr = 0;
    {
        double sum_info = 0.0;
        std::vector<int32_t> reco_cnt((size_t)I, 0);
        for (int32_t u = 0; u < U; ++u) {
            const int topk = topn <= rec_counts[u] ? topn : rec_counts[u];
            for (int i = 0; i < topk; ++i) {
                const int32_t it = rec_items[(int64_t)u * topn + i];
                reco_cnt[(size_t)it]++;
                const int32_t c = item_purchase_num[it];
                if (c > 0) sum_info += -log(((double)c) / U);
            }
        }
        out[6] = sum_info / (U * log(2.0));
        double sum_ent = 0.0;
        for (int32_t i = 0; i < I; ++i) if (reco_cnt[(size_t)i] > 0) { const double p = ((double)reco_cnt[(size_t)i]) / U; sum_ent += p * (-log(p)); }
        out[7] = sum_ent / log(2.0);
    }
    int64_t nz = 0, nz_ap = 0;
    std::vector<std::pair<uint64_t, int32_t>> order;
    std::vector<double> hitvals;
    for (int32_t u = 0; u < U; ++u) {
        const int64_t tb = t_rowptr[u], te = t_rowptr[u + 1];
        const int64_t nt = te - tb;
        if (nt <= 0) continue;
        ++nz;
        const int32_t* r = rec_items + (int64_t)u * topn;
        const int topk = topn <= rec_counts[u] ? topn : rec_counts[u];
        auto in_test = [&](int32_t key, double* val) {
            const int32_t* p = std::lower_bound(t_col + tb, t_col + te, key);
            if (p != t_col + te && *p == key) { if (val) *val = t_val[p - t_col]; return true; }
            return false;
        };
        // Precision / Recall / AP / RR / DCG in list order
        int hits = 0; double tmp_prec = 0.0, dcg = 0.0; bool first = true, has_dcg = false;
        hitvals.clear();
        for (int i = 0; i < topk; ++i) {
            double v = 0.0;
            if (in_test(r[i], &v)) {
                ++hits;
                tmp_prec += 1.0 * hits / (i + 1);
                if (first) { rr += 1.0 / (i + 1.0); first = false; }
                has_dcg = true;
                dcg += v / (log((double)(i + 2)) / log(2.0));           // Maths.log(n, 2) = Math.log(n) / Math.log(2)
                hitvals.push_back(v);
            }
        }
        prec += hits / (topn + 0.0);
        rec += hits / (nt + 0.0);
        if (topk != 0) { ap += tmp_prec / (double)(nt < topk ? nt : topk); ++nz_ap; }
        if (has_dcg && dcg != 0.0) {
            std::sort(hitvals.begin(), hitvals.end(), [](double a, double b) { return a > b; });
            double idcg = 0.0;
            for (size_t i = 0; i < hitvals.size(); ++i) idcg += hitvals[i] / (log((double)(i + 2)) / log(2.0));
            if (idcg != 0.0) ndcg += dcg / idcg;
        }
        // AUC
        {
            const int num_dropped_items = num_dropped[u] - topk;
            int rel = 0, miss = 0;
            for (int i = 0; i < topk; ++i) { if (in_test(r[i], nullptr)) ++rel; else ++miss; }    // list keys are distinct
            const int64_t n_items = (int64_t)num_dropped_items + topk;
            const int64_t n_pairs = (n_items - rel) * rel;
            if (n_pairs == 0) { auc += 0.5; continue; }
            const uint32_t cap = jhashset_capacity(nt);
            order.clear();
            for (int64_t e = tb; e < te; ++e) order.push_back({((uint64_t)jhash_bucket(t_col[e], cap) << 32) | (uint32_t)(e - tb), t_col[e]});
            std::sort(order.begin(), order.end());
            int64_t correct = 0; int h2 = 0;
            for (auto& o : order) {
                bool in_rec = false;
                for (int i = 0; i < topk; ++i) if (r[i] == o.second) { in_rec = true; break; }
                if (!in_rec) correct += h2; else ++h2;
            }
            correct += (int64_t)h2 * (num_dropped_items - miss);
            auc += (correct + 0.0) / (double)n_pairs;
        }
    }
    out[0] = nz > 0 ? auc / nz : 0.0;
    out[1] = nz_ap > 0 ? ap / nz_ap : 0.0;
    out[2] = nz > 0 ? ndcg / nz : 0.0;
    out[3] = nz > 0 ? prec / nz : 0.0;
    out[4] = nz > 0 ? rec / nz : 0.0;
    out[5] = nz > 0 ? rr / nz : 0.0;
}

// The three ranking evaluators outside the default list (reachable through rec.eval.classes): out[0] HitRate
// (eval/ranking/HitRateEvaluator.java:33-62; leave-one-out only -- NaN here when a user has more than one test item, where the
// reference throws), out[1] ARHR (AverageReciprocalHitRankEvaluator.java:33-56: reciprocal rank of the FIRST test item of the
// user, i.e. its lowest item id), out[2] IDCG (IdealDCGEvaluator.java:34-52: sum_{i < |test|} 1 / log2(i + 2), averaged).
LRO_API void lro_eval_ranking_extra(int32_t U, int32_t topn, const int32_t* rec_items, const int32_t* rec_counts,
                                    const int64_t* t_rowptr, const int32_t* t_col, double* out) {
    int64_t hits = 0, nz_hit = 0, nz = 0;
    bool loo = true;
    double arhr = 0.0, idcg_sum = 0.0;
    for (int32_t u = 0; u < U; ++u) {
        const int64_t tb = t_rowptr[u], nt = t_rowptr[u + 1] - tb;
        if (nt <= 0) continue;
        ++nz;
        const int32_t* r = rec_items + (int64_t)u * topn;
        const int topk = topn <= rec_counts[u] ? topn : rec_counts[u];
        if (nt == 1) {
            for (int i = 0; i < topk; ++i) if (r[i] == t_col[tb]) { ++hits; break; }
            ++nz_hit;
        } else loo = false;
        for (int i = 0; i < topk; ++i) if (r[i] == t_col[tb]) { arhr += 1.0 / (i + 1.0); break; }
        double idcg = 0.0;
        for (int64_t i = 0; i < nt; ++i) idcg += 1 / (log((double)i + 2.0) / log(2.0));
        idcg_sum += idcg;
    }
    out[0] = !loo ? std::nan("") : (nz_hit > 0 ? 1.0 * hits / nz_hit : 0.0);
    out[1] = nz > 0 ? arhr / nz : 0.0;
    out[2] = nz > 0 ? idcg_sum / nz : 0.0;
}

LRO_API void lro_predict_pairs(int32_t model, int32_t k, const double* P, const double* Q, const double* bu,
                               const double* bi, double mu, const int32_t* us, const int32_t* is, int64_t n, double* out) {
    for (int64_t t = 0; t < n; ++t) out[t] = predict_raw(model, k, P, Q, bu, bi, mu, us[t], is[t]);
}

// -------------------------------------------------------------------------------------
// Top-N: recommender/MatrixRecommender.java:153-201 + util/Lists.java:416-468
// (+ recommender/item/RecommendedList.java:85-88).
// java.util.PriorityQueue (JDK 8) siftUp / siftDown restated; comparator for inverse=true is
// Double.compareTo ascending (min-heap); entry replaces the min only if STRICTLY greater;
// result = heap-array order, then a stable descending sort (Collections.sort == TimSort).
// -------------------------------------------------------------------------------------
static inline int jcompare(double a, double b) {     // Double.compare
    if (a < b) return -1;
    if (a > b) return 1;
    int64_t x, y; memcpy(&x, &a, 8); memcpy(&y, &b, 8);   // doubleToLongBits (NaN never reaches here)
    return x == y ? 0 : (x < y ? -1 : 1);
}
struct KV { int32_t key; double value; };
struct JHeap {
    std::vector<KV> q; int size = 0;
    void sift_up(int kpos, KV x) {
        while (kpos > 0) {
            int parent = (int)((unsigned)(kpos - 1) >> 1);
            if (jcompare(x.value, q[parent].value) >= 0) break;
            q[kpos] = q[parent]; kpos = parent;
        }
        q[kpos] = x;
    }
    void sift_down(int kpos, KV x) {
        int half = (int)((unsigned)size >> 1);
        while (kpos < half) {
            int child = (kpos << 1) + 1; int right = child + 1;
            if (right < size && jcompare(q[child].value, q[right].value) > 0) child = right;
            if (jcompare(x.value, q[child].value) <= 0) break;
            q[kpos] = q[child]; kpos = child;
        }
        q[kpos] = x;
    }
    void add(KV x) { int i = size; size = i + 1; if (i == 0) q[0] = x; else sift_up(i, x); }
    void poll() { int s = --size; KV x = q[s]; if (s != 0) sift_down(0, x); }
};

static int topn_user(int model, int32_t I, int32_t k, const double* P, const double* Q, const double* bu,
                     const double* bi, double mu, const int32_t* tcol, int64_t tlen, int32_t u, int32_t topN,
                     int32_t* out_items, double* out_scores, JHeap& h, std::vector<KV>& list) {
    // MatrixRecommender.java:168-190 : all items not in the train row, NaN dropped
    list.clear();
    int64_t tp = 0;
    for (int32_t i = 0; i < I; ++i) {
        if (tp < tlen && tcol[tp] == i) { ++tp; continue; }
        double pr = predict_raw(model, k, P, Q, bu, bi, mu, u, i);
        if (std::isnan(pr)) continue;
        list.push_back(KV{i, pr});
    }
    // Lists.sortKeyValueListTopK(list, true, topN) : Lists.java:416-445
    int kk = (int)list.size() > topN ? topN : (int)list.size();
    if (kk == 0) return 0;
    h.q.resize(kk); h.size = 0;
    size_t it = 0;
    for (int t = 0; t < kk; ++t) h.add(list[it++]);
    for (; it < list.size(); ++it) {
        int res = jcompare(list[it].value, h.q[0].value);
        if (-res < 0) { h.poll(); h.add(list[it]); }
    }
    std::vector<KV> out(h.q.begin(), h.q.begin() + h.size);
    std::stable_sort(out.begin(), out.end(), [](const KV& a, const KV& b) { return jcompare(a.value, b.value) > 0; });
    for (int t = 0; t < kk; ++t) { out_items[t] = out[t].key; out_scores[t] = out[t].value; }
    return kk;
}

// users == NULL -> users 0..nq-1 (MatrixRecommender.recommendRank() :137-144).
// nthreads mirrors contextList.parallelStream() (:164); results are per-user independent.
LRO_API void lro_recommend_rank(int32_t model, int32_t U, int32_t I, int32_t k, const double* P, const double* Q,
                                const double* bu, const double* bi, double mu, const int64_t* tr_rowptr,
                                const int32_t* tr_col, int32_t topN, const int32_t* users, int32_t nq,
                                int32_t* out_items, double* out_scores, int32_t* out_counts, int32_t nthreads) {
    (void)U;
#ifdef _OPENMP
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
#endif
    {
        JHeap h; std::vector<KV> list; list.reserve((size_t)I);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (int32_t c = 0; c < nq; ++c) {
            int32_t u = users ? users[c] : c;
            const int64_t b = tr_rowptr ? tr_rowptr[u] : 0, e = tr_rowptr ? tr_rowptr[u + 1] : 0;
            for (int t = 0; t < topN; ++t) { out_items[(int64_t)c * topN + t] = -1; out_scores[(int64_t)c * topN + t] = 0.0; }
            out_counts[c] = topn_user(model, I, k, P, Q, bu, bi, mu, tr_col ? tr_col + b : nullptr, e - b, u, topN,
                                      out_items + (int64_t)c * topN, out_scores + (int64_t)c * topN, h, list);
        }
    }
}

// Raw heap trace for the PriorityQueue pin test: feeds `n` values through
// sortKeyValueListTopK(inverse=true,k) and returns the heap-array order BEFORE the final sort.
LRO_API int32_t lro_heap_trace(const double* values, int32_t n, int32_t k, int32_t* heap_keys) {
    int kk = n > k ? k : n;
    if (kk == 0) return 0;
    JHeap h; h.q.resize(kk);
    int it = 0;
    for (int t = 0; t < kk; ++t, ++it) h.add(KV{it, values[it]});
    for (; it < n; ++it) if (-jcompare(values[it], h.q[0].value) < 0) { h.poll(); h.add(KV{it, values[it]}); }
    for (int t = 0; t < kk; ++t) heap_keys[t] = h.q[t].key;
    return kk;
}

// -------------------------------------------------------------------------------------
// "Best-effort CPU" variant (NOT the reference's behaviour, reported separately by bench.py):
// fp32 Hogwild over an explicit order with OpenMP threads.  SURVEY.md 8(d).
// -------------------------------------------------------------------------------------
LRO_API double lro_sgd_epoch_hogwild_f32(int32_t model, const int32_t* us, const int32_t* is, const float* rs, int64_t n,
                                         int32_t k, float* P, float* Q, float* bu, float* bi, float mu,
                                         float lr, float regU, float regI, float regB, int32_t nthreads) {
    double loss = 0.0;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) reduction(+ : loss) schedule(static)
#endif
    for (int64_t t = 0; t < n; ++t) {
        const int32_t u = us[t], i = is[t];
        float* pu = P + (int64_t)u * k; float* qi = Q + (int64_t)i * k;
        float d = 0.f;
        for (int f = 0; f < k; ++f) d += pu[f] * qi[f];
        float l = 0.f;
        if (model == LRO_BIASEDMF) {
            d += bu[u] + bi[i] + mu;
            const float e = rs[t] - d;
            const float ub = bu[u], ib = bi[i];
            bu[u] = ub + lr * (e - regB * ub); bi[i] = ib + lr * (e - regB * ib);
            l = e * e + regB * ub * ub + regB * ib * ib;
            for (int f = 0; f < k; ++f) {
                const float uf = pu[f], itf = qi[f];
                pu[f] = uf + lr * (e * itf - regU * uf); qi[f] = itf + lr * (e * uf - regI * itf);
                l += regU * uf * uf + regI * itf * itf;
            }
        } else {
            const float e = rs[t] - d;
            l = e * e;
            for (int f = 0; f < k; ++f) {
                const float uf = pu[f], itf = qi[f];
                pu[f] = uf + lr * (e * itf - regU * uf); qi[f] = itf + lr * (e * uf - regI * itf);
                l += regU * uf * uf + regI * itf * itf;
            }
        }
        loss += (double)l;
    }
    return 0.5 * loss;
}

LRO_API int32_t lro_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
LRO_API const char* lro_version() { return "lrk-oracle 0.1 (LibRec 3.0.0 MF path restatement; parity unpinned)"; }
