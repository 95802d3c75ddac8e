// NOT COMPILED HERE.  rec.recommender.class=net.librec.recommender.cuda.SVDPlusPlusCudaRecommender
// Replaces the trainModel() of recommender/cf/rating/SVDPlusPlusRecommender.java:62-109; single GPU.  The implicit-feedback factors
// (impItemFactors, :55-56) travel through lrk_set_matrix / lrk_get_matrix("svdpp.y"), rec.impItem.regularization through
// lrk_set_param.  predict() stays the reference's (:126-148: it needs the per-user sum of the implicit factors over the train row)
// and reads the matrices trainModel() copied back, so recommendRating and every evaluator work unchanged; recommendRank is
// inherited from MatrixRecommender (the native top-N does not cover this model).
package net.librec.recommender.cuda;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.DoubleBuffer;
import java.nio.charset.StandardCharsets;

import net.librec.common.LibrecException;
import net.librec.data.structure.AbstractBaseDataEntry;
import net.librec.data.structure.LibrecDataList;
import net.librec.math.structure.DenseMatrix;
import net.librec.math.structure.SequentialSparseVector;
import net.librec.math.structure.Vector;
import net.librec.recommender.item.RecommendedList;

public class SVDPlusPlusCudaRecommender extends BiasedMFCudaRecommender {
    protected DenseMatrix impItemFactors;
    protected double regImpItem;
    private static final byte[] Y = "svdpp.y".getBytes(StandardCharsets.UTF_8);

    @Override int model() { return LibrecB200.MODEL_SVDPP; }

    @Override
    protected void setup() throws LibrecException {
        super.setup();                                                     // BiasedMF setup: biases after the factor matrices
        regImpItem = conf.getDouble("rec.impItem.regularization", 0.015d);  // SVDPlusPlusRecommender.java:52
        impItemFactors = new DenseMatrix(numItems, numFactors);             // :55-56, same RNG order as the reference
        impItemFactors.init(initMean, initStd);
        check(LibrecB200.setParam(handle, "svdpp.reg_imp".getBytes(StandardCharsets.UTF_8), regImpItem));
    }

    @Override
    protected void afterSetFactors() throws LibrecException {              // hook of MatrixFactorizationCudaRecommender.trainModel()
        ByteBuffer y = LibrecB200.hostAlloc(8L * numItems * numFactors).order(ByteOrder.nativeOrder());
        DoubleBuffer d = y.asDoubleBuffer();
        for (int r = 0; r < numItems; r++) { d.position(r * numFactors); d.put(impItemFactors.getValues()[r]); }
        check(LibrecB200.setMatrix(handle, Y, y));
        LibrecB200.hostFree(y);
    }

    @Override
    protected void afterGetFactors() throws LibrecException {
        ByteBuffer y = LibrecB200.hostAlloc(8L * numItems * numFactors).order(ByteOrder.nativeOrder());
        check(LibrecB200.getMatrix(handle, Y, y));
        DoubleBuffer d = y.asDoubleBuffer();
        for (int r = 0; r < numItems; r++) { d.position(r * numFactors); d.get(impItemFactors.getValues()[r]); }
        LibrecB200.hostFree(y);
    }

    @Override
    protected double predict(int userIndex, int itemIndex) throws LibrecException {    // SVDPlusPlusRecommender.java:137-148
        SequentialSparseVector userVector = trainMatrix.row(userIndex);
        double[] fv = new double[numFactors];
        for (Vector.VectorEntry ve : userVector)
            for (int f = 0; f < numFactors; f++) fv[f] = impItemFactors.get(ve.index(), f) + fv[f];
        double scale = userVector.getNumEntries() > 0 ? Math.pow(userVector.getNumEntries(), -0.5) : 0.0;
        double value = userBiases.get(userIndex) + itemBiases.get(itemIndex) + globalMean;
        for (int f = 0; f < numFactors; f++)
            value += (fv[f] * scale + userFactors.get(userIndex, f)) * itemFactors.get(itemIndex, f);
        return value;
    }

    @Override
    public RecommendedList recommendRank(LibrecDataList<AbstractBaseDataEntry> dataList) throws LibrecException {
        return recommendRankReference(dataList);                           // MatrixRecommender.java:153-201 on predict() above
    }
}
