// NOT COMPILED HERE.  rec.recommender.class=net.librec.recommender.cuda.BPRCudaRecommender
package net.librec.recommender.cuda;

public class BPRCudaRecommender extends MatrixFactorizationCudaRecommender {
    @Override int model() { return LibrecB200.MODEL_BPR; }
}
