// C entry points over the C++ host mirror, so that the Python tests (and any other FFI) can drive
// RecommenderJob / the recommender classes the way the reference's JUnit cases do
// (core/src/test/java/net/librec/recommender/cf/rating/BiasedMFTestCase.java:50-56:
//  conf.addResource(...); new RecommenderJob(conf).runJob()).
#include "librec_host.hpp"

#include <cstring>
#include <string>

using namespace librec;

namespace {
thread_local std::string g_err;
struct Job {
    std::unique_ptr<RecommenderJob> job;
    std::string log_text;
};
SequentialAccessSparseMatrix make_csr(int32_t rows, int32_t cols, const int64_t* rowptr, const int32_t* col, const double* val) {
    SequentialAccessSparseMatrix m;
    m.numRows = rows; m.numCols = cols;
    m.rowptr.assign(rowptr, rowptr + rows + 1);
    const int64_t n = rowptr[rows];
    m.col.assign(col, col + n);
    m.val.assign(val, val + n);
    return m;
}
}  // namespace

#define LRH_API extern "C" __attribute__((visibility("default")))

LRH_API const char* lrh_last_error() { return g_err.c_str(); }

LRH_API void* lrh_job_create(const char* properties_text) {
    try {
        Configuration conf;
        conf.load_properties(properties_text ? properties_text : "");
        Job* j = new Job();
        j->job.reset(new RecommenderJob(conf));       // seeds Randoms like RecommenderJob.java:72-79
        return j;
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
LRH_API void lrh_job_destroy(void* h) { delete (Job*)h; }

LRH_API int lrh_job_set_data(void* h, int32_t U, int32_t I, const int64_t* tr_rowptr, const int32_t* tr_col, const double* tr_val,
                             const int64_t* te_rowptr, const int32_t* te_col, const double* te_val) {
    try {
        Job* j = (Job*)h;
        j->job->setData(make_csr(U, I, tr_rowptr, tr_col, tr_val), make_csr(U, I, te_rowptr, te_col, te_val));
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// 0 ok; -1 LibrecException; -2 other exception (IndexOutOfBounds etc.)
LRH_API int lrh_job_run(void* h) {
    Job* j = (Job*)h;
    try {
        j->job->runJob();
        return 0;
    } catch (const LibrecException& e) { g_err = e.what(); return -1; }
    catch (const std::exception& e) { g_err = e.what(); return -2; }
}
LRH_API double lrh_job_metric(void* h, const char* name) {
    Job* j = (Job*)h;
    auto it = j->job->evaluatedMap.find(name);
    return it == j->job->evaluatedMap.end() ? -1.0 : it->second;
}
LRH_API const char* lrh_job_log(void* h) {
    Job* j = (Job*)h;
    j->log_text.clear();
    const auto& src = j->job->log.empty() && j->job->recommender ? j->job->recommender->log() : j->job->log;
    for (const auto& l : src) { j->log_text += l; j->log_text += '\n'; }
    return j->log_text.c_str();
}
// factors after training (any pointer may be NULL)
LRH_API int lrh_job_factors(void* h, double* P, double* Q, double* bu, double* bi, double* mu) {
    Job* j = (Job*)h;
    if (!j->job->recommender) { g_err = "job has not run"; return -1; }
    auto& r = *j->job->recommender;
    if (P) memcpy(P, r.getUserFactors().values.data(), r.getUserFactors().values.size() * 8);
    if (Q) memcpy(Q, r.getItemFactors().values.data(), r.getItemFactors().values.size() * 8);
    if (bu && !r.getUserBiases().values.empty()) memcpy(bu, r.getUserBiases().values.data(), r.getUserBiases().values.size() * 8);
    if (bi && !r.getItemBiases().values.empty()) memcpy(bi, r.getItemBiases().values.data(), r.getItemBiases().values.size() * 8);
    if (mu) *mu = r.getGlobalMean();
    return 0;
}
// the RecommendedList of the last run flattened: counts[ctx], then (key,value) pairs in list order
LRH_API int64_t lrh_job_list_size(void* h, int32_t* num_contexts) {
    Job* j = (Job*)h;
    int64_t n = 0;
    for (const auto& l : j->job->recommendedList.lists) n += (int64_t)l.size();
    if (num_contexts) *num_contexts = j->job->recommendedList.size();
    return n;
}
LRH_API void lrh_job_list_copy(void* h, int32_t* counts, int32_t* keys, double* values) {
    Job* j = (Job*)h;
    int64_t t = 0;
    int c = 0;
    for (const auto& l : j->job->recommendedList.lists) {
        counts[c++] = (int32_t)l.size();
        for (const auto& kv : l) { keys[t] = kv.key; values[t] = kv.value; ++t; }
    }
}
// host-logic probes used by the CPU tests (no GPU needed)
// the four default rating measures (eval/Measure.java:100-107) over a flat test CSR and the predictions in the same order:
// out = RMSE, MSE, MAE, MPE
LRH_API int lrh_probe_rating_measures(int32_t U, int32_t I, const int64_t* rowptr, const int32_t* col, const double* val,
                                      const double* pred, double mpe, double* out) {
    try {
        SequentialAccessSparseMatrix t;
        t.numRows = U; t.numCols = I;
        t.rowptr.assign(rowptr, rowptr + U + 1);
        t.col.assign(col, col + rowptr[U]); t.val.assign(val, val + rowptr[U]);
        RecommendedList rec;
        for (int u = 0; u < U; ++u) {
            rec.addList();
            for (int64_t e = rowptr[u]; e < rowptr[u + 1]; ++e) rec.add(u, col[e], pred[e]);
        }
        out[0] = evaluateRMSE(t, rec); out[1] = evaluateMSE(t, rec); out[2] = evaluateMAE(t, rec); out[3] = evaluateMPE(t, rec, mpe);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// the eight default ranking measures over flat CSR train / test and padded lists [U x topn] (counts[u] valid entries):
// out = AUC, AP, NDCG, PRECISION, RECALL, RR, Novelty, Entropy
LRH_API int lrh_probe_ranking_measures(int32_t U, int32_t I, const int64_t* tr_rowptr, const int32_t* tr_col, const int64_t* te_rowptr,
                                       const int32_t* te_col, const double* te_val, int32_t topn, const int32_t* items,
                                       const int32_t* counts, double* out) {
    try {
        SequentialAccessSparseMatrix tr, te;
        tr.numRows = te.numRows = U; tr.numCols = te.numCols = I;
        tr.rowptr.assign(tr_rowptr, tr_rowptr + U + 1); tr.col.assign(tr_col, tr_col + tr_rowptr[U]); tr.val.assign((size_t)tr_rowptr[U], 1.0);
        te.rowptr.assign(te_rowptr, te_rowptr + U + 1); te.col.assign(te_col, te_col + te_rowptr[U]); te.val.assign(te_val, te_val + te_rowptr[U]);
        RecommendedList rec;
        for (int u = 0; u < U; ++u) {
            rec.addList();
            for (int t = 0; t < counts[u]; ++t) rec.add(u, items[(int64_t)u * topn + t], 0.0);
        }
        std::map<std::string, double> m;
        evaluateRanking(tr, te, rec, topn, &m);
        static const char* names[8] = {"AUC", "AP", "NDCG", "PRECISION", "RECALL", "RR", "Novelty", "Entropy"};
        for (int i = 0; i < 8; ++i) out[i] = m[std::string(names[i]) + " top " + std::to_string(topn)];
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// HitRate / ARHR / IDCG over a flat test CSR and padded lists: out[3]; returns -1 (message in lrh_last_error) when HitRate is asked
// for data that is not leave-one-out
LRH_API int lrh_probe_ranking_extra(int32_t U, int32_t I, const int64_t* te_rowptr, const int32_t* te_col, int32_t topn,
                                    const int32_t* items, const int32_t* counts, int want_hitrate, double* out) {
    try {
        SequentialAccessSparseMatrix te;
        te.numRows = U; te.numCols = I;
        te.rowptr.assign(te_rowptr, te_rowptr + U + 1); te.col.assign(te_col, te_col + te_rowptr[U]); te.val.assign((size_t)te_rowptr[U], 1.0);
        RecommendedList rec;
        for (int u = 0; u < U; ++u) {
            rec.addList();
            for (int t = 0; t < counts[u]; ++t) rec.add(u, items[(int64_t)u * topn + t], 0.0);
        }
        std::map<std::string, double> m;
        evaluateRankingExtra(te, rec, topn, want_hitrate != 0, &m);
        const std::string sfx = " top " + std::to_string(topn);
        out[0] = want_hitrate ? m["HitRate" + sfx] : 0.0; out[1] = m["ARHR" + sfx]; out[2] = m["IDCG" + sfx];
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// ---- data model alone (no GPU needed): properties -> TextDataModel.buildDataModel() -> flat CSR arrays
struct DataModelBox { std::unique_ptr<TextDataModel> dm; std::string tmp; };
LRH_API void* lrh_datamodel_build(const char* properties_text) {
    try {
        Configuration conf;
        conf.load_properties(properties_text ? properties_text : "");
        if (conf.has("rec.random.seed")) Randoms::seed(conf.getLong("rec.random.seed", 1));      // RecommenderJob.java:74-77
        DataModelBox* b = new DataModelBox();
        b->dm.reset(new TextDataModel(conf));
        b->dm->buildDataModel();
        return b;
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
LRH_API void lrh_datamodel_destroy(void* h) { delete (DataModelBox*)h; }
// AbstractDataModel.hasNextFold(): 1 = the train / test matrices now hold the next fold, 0 = no fold left
LRH_API int lrh_datamodel_next_fold(void* h) {
    try { return ((DataModelBox*)h)->dm->hasNextFold() ? 1 : 0; } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
LRH_API int lrh_datamodel_num_folds(void* h) { return ((DataModelBox*)h)->dm->numFolds; }
static const SequentialAccessSparseMatrix& dm_matrix(void* h, int which) {
    TextDataModel* d = ((DataModelBox*)h)->dm.get();
    return which == 0 ? d->preference : (which == 1 ? d->train : (which == 2 ? d->test : d->valid));
}
LRH_API void lrh_datamodel_dims(void* h, int which, int32_t* U, int32_t* I, int64_t* nnz) {
    const SequentialAccessSparseMatrix& m = dm_matrix(h, which);
    *U = m.numRows; *I = m.numCols; *nnz = m.size();
}
LRH_API void lrh_datamodel_copy(void* h, int which, int64_t* rowptr, int32_t* col, double* val) {
    const SequentialAccessSparseMatrix& m = dm_matrix(h, which);
    memcpy(rowptr, m.rowptr.data(), m.rowptr.size() * sizeof(int64_t));
    memcpy(col, m.col.data(), m.col.size() * sizeof(int32_t));
    memcpy(val, m.val.data(), m.val.size() * sizeof(double));
}
LRH_API const char* lrh_datamodel_raw_id(void* h, int is_item, int32_t inner) {
    DataModelBox* b = (DataModelBox*)h;
    const auto& ids = is_item ? b->dm->itemIds : b->dm->userIds;
    b->tmp = (inner >= 0 && (size_t)inner < ids.size()) ? ids[(size_t)inner] : std::string();
    return b->tmp.c_str();
}
// job/RecommenderJob.java:281-306; returns the path written ("" when the list is empty), NULL on error
LRH_API const char* lrh_job_save_result(void* h) {
    Job* j = (Job*)h;
    try { j->log_text = j->job->saveResult(); return j->log_text.c_str(); }
    catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}

LRH_API void lrh_randoms_seed(long long s) { Randoms::seed(s); }
LRH_API int lrh_randoms_uniform_int(int range) { return Randoms::uniform(range); }
LRH_API double lrh_randoms_uniform() { return Randoms::uniform(); }
LRH_API double lrh_randoms_gaussian(double mu, double sigma) { return Randoms::gaussian(mu, sigma); }
LRH_API void lrh_format_double(double v, char* out, int cap) { std::string s = java_double_to_string(v); strncpy(out, s.c_str(), cap - 1); out[cap - 1] = 0; }
LRH_API void lrh_format_float(float v, char* out, int cap) { std::string s = java_float_to_string(v); strncpy(out, s.c_str(), cap - 1); out[cap - 1] = 0; }
LRH_API double lrh_conf_probe(const char* props, const char* key, int kind, double def) {
    Configuration c; c.load_properties(props);
    switch (kind) {
        case 0: return c.getInt(key, (int)def);
        case 1: return (double)c.getFloat(key, (float)def);
        case 2: return c.getDouble(key, def);
        case 3: return c.getBoolean(key, def != 0) ? 1.0 : 0.0;
        default: return c.has(key) ? 1.0 : 0.0;
    }
}
