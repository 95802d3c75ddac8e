// NOT COMPILED HERE.  rec.recommender.class=net.librec.recommender.cuda.GBPRCudaRecommender
// Replaces recommender/cf/ranking/GBPRRecommender.java:82-172 (group-preference BPR); single GPU.  Item biases only: the shim passes
// null user biases to lrk_set_factors; rec.gpbr.rho / rec.gpbr.gsize go through lrk_set_param, rec.bias.regularization is the reg_b
// argument of lrk_sgd_epoch.  (The Guava row / column caches of the reference's setup() are not needed: the library builds the
// train CSC on the device.)
package net.librec.recommender.cuda;

import java.nio.charset.StandardCharsets;

import net.librec.common.LibrecException;
import net.librec.math.structure.VectorBasedDenseVector;

public class GBPRCudaRecommender extends MatrixFactorizationCudaRecommender {
    @Override int model() { return LibrecB200.MODEL_GBPR; }

    @Override
    protected void setup() throws LibrecException {
        super.setup();
        itemBiases = new VectorBasedDenseVector(numItems);             // GBPRRecommender.java:68-69 (init() = zeros)
        itemBiases.init();
        regBias = conf.getDouble("rec.bias.regularization", 0.01);     // :73
        check(LibrecB200.setParam(handle, "gbpr.rho".getBytes(StandardCharsets.UTF_8), conf.getFloat("rec.gpbr.rho", 1.5f)));
        check(LibrecB200.setParam(handle, "gbpr.gsize".getBytes(StandardCharsets.UTF_8), conf.getInt("rec.gpbr.gsize", 2)));
    }

    @Override
    protected double predict(int userIdx, int itemIdx) throws LibrecException {   // GBPRRecommender.java:194-196
        return itemBiases.get(itemIdx) + super.predict(userIdx, itemIdx);
    }
}
