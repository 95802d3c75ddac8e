// NOT COMPILED HERE.  rec.recommender.class=net.librec.recommender.cuda.PMFCudaRecommender
// Vanilla PMF (PMFSimilarityRecommender.java:59-90 loop with numIterations), not the fork's tag variant.
package net.librec.recommender.cuda;

public class PMFCudaRecommender extends MatrixFactorizationCudaRecommender {
    @Override int model() { return LibrecB200.MODEL_PMF; }
}
