// NOT COMPILED HERE.  rec.recommender.class=net.librec.recommender.cuda.RankSGDCudaRecommender
// Replaces recommender/cf/ranking/RankSGDRecommender.java:62-108 (one update per train entry against a negative item drawn
// by popularity); single GPU.  The item probability list of the reference's setup() (:42-58) is built on the device from
// the train CSR, so nothing but the model id differs from the other shims.
package net.librec.recommender.cuda;

public class RankSGDCudaRecommender extends MatrixFactorizationCudaRecommender {
    @Override int model() { return LibrecB200.MODEL_RANKSGD; }
}
