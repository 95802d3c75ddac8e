// NOT COMPILED HERE.  rec.recommender.class=net.librec.recommender.cuda.BiasedMFCudaRecommender
package net.librec.recommender.cuda;

import net.librec.common.LibrecException;
import net.librec.math.structure.VectorBasedDenseVector;

public class BiasedMFCudaRecommender extends MatrixFactorizationCudaRecommender {
    @Override int model() { return LibrecB200.MODEL_BIASEDMF; }

    @Override
    protected void setup() throws LibrecException {
        super.setup();
        regBias = conf.getDouble("rec.bias.regularization", 0.01);     // BiasedMFRecommender.java:56
        userBiases = new VectorBasedDenseVector(numUsers);
        itemBiases = new VectorBasedDenseVector(numItems);
        userBiases.init(initMean, initStd);                            // same RNG order as BiasedMFRecommender.java:59-63
        itemBiases.init(initMean, initStd);
    }

    @Override
    protected double predict(int userIdx, int itemIdx) throws LibrecException {   // BiasedMFRecommender.java:118-120
        return userFactors.row(userIdx).dot(itemFactors.row(itemIdx)) + userBiases.get(userIdx) + itemBiases.get(itemIdx) + globalMean;
    }
}
