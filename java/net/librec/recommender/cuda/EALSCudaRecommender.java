// NOT COMPILED HERE.  rec.recommender.class=net.librec.recommender.cuda.EALSCudaRecommender
// Replaces the trainModel() of recommender/cf/ranking/EALSRecommender.java:114-214 (element-wise ALS); single GPU.  setup() keeps
// the reference's confidences and weightMatrix() (:53-112) in Java; the confidences travel through lrk_set_matrix("eals.confidences").
// The native iteration is bit-identical to the reference's (fp64, the reference's operation order).
package net.librec.recommender.cuda;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.charset.StandardCharsets;

import it.unimi.dsi.fastutil.doubles.Double2DoubleOpenHashMap;

import net.librec.common.LibrecException;
import net.librec.math.structure.DenseMatrix;
import net.librec.math.structure.MatrixEntry;

public class EALSCudaRecommender extends MatrixFactorizationCudaRecommender {
    protected float weightCoefficient, ratio, overallWeight;
    protected int WRMFJudge;
    protected double[] confidences;

    @Override int model() { return LibrecB200.MODEL_EALS; }

    public double weight(double value) {                                   // EALSRecommender.java:85-95
        return (WRMFJudge == 1 || WRMFJudge == 2) ? 1.0 + weightCoefficient * value : 1.0;
    }

    @Override
    protected void beforeStage() throws LibrecException {                  // EALSRecommender.java:53-112
        weightCoefficient = conf.getFloat("rec.wrmf.weight.coefficient", 4.0f);
        ratio = conf.getFloat("rec.eals.ratio", 0.4f);
        overallWeight = conf.getFloat("rec.eals.overall", 128.0f);
        WRMFJudge = conf.getInt("rec.eals.wrmf.judge", 1);
        confidences = new double[numItems];
        if (WRMFJudge == 0 || WRMFJudge == 2) {
            double sumPopularity = 0.0;
            for (int itemIdx = 0; itemIdx < numItems; itemIdx++) {
                double alphaPopularity = Math.pow(trainMatrix.column(itemIdx).getNumEntries() * 1.0 / numRates, ratio);
                confidences[itemIdx] = overallWeight * alphaPopularity;
                sumPopularity += alphaPopularity;
            }
            for (int itemIdx = 0; itemIdx < numItems; itemIdx++) confidences[itemIdx] = confidences[itemIdx] / sumPopularity;
        } else
            for (int itemIdx = 0; itemIdx < numItems; itemIdx++) confidences[itemIdx] = 1;
        Double2DoubleOpenHashMap ratingWeightMap = new Double2DoubleOpenHashMap();
        for (double rating : ratingScale) ratingWeightMap.putIfAbsent(rating, weight(rating));
        for (MatrixEntry matrixEntry : trainMatrix) matrixEntry.set(ratingWeightMap.get(matrixEntry.get()));
    }

    @Override
    protected void afterStage() throws LibrecException {
        ByteBuffer c = LibrecB200.hostAlloc(8L * numItems).order(ByteOrder.nativeOrder());
        c.asDoubleBuffer().put(confidences);
        check(LibrecB200.setMatrix(handle, "eals.confidences".getBytes(StandardCharsets.UTF_8), c));
        LibrecB200.hostFree(c);
    }

    @Override
    protected void beforeTrainModel() throws LibrecException {
        userFactors = new DenseMatrix(numUsers, numFactors);               // EALSRecommender.java:125
    }
}
