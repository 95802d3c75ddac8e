// NOT COMPILED HERE.  rec.recommender.class=net.librec.recommender.cuda.AoBPRCudaRecommender
// Replaces recommender/cf/ranking/AoBPRRecommender.java:82-172 (BPR with adaptive oversampling of the negative item); single GPU.
// rec.item.distribution.parameter goes through lrk_set_param; the rank distribution, the per-factor item rankings and their refresh
// every |I| ln |I| samples (:60-78,186-200) live in the library (csrc/sgd_aobpr.cuh).
package net.librec.recommender.cuda;

import java.nio.charset.StandardCharsets;

import net.librec.common.LibrecException;

public class AoBPRCudaRecommender extends MatrixFactorizationCudaRecommender {
    @Override int model() { return LibrecB200.MODEL_AOBPR; }

    @Override
    protected void setup() throws LibrecException {
        super.setup();
        check(LibrecB200.setParam(handle, "aobpr.lambda".getBytes(StandardCharsets.UTF_8),
                                  conf.getFloat("rec.item.distribution.parameter")));     // AoBPRRecommender.java:63 (no default)
    }
}
