// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no JDK).  Drop-in for the MF hot path of LibRec 3.0.0:
// subclasses only override what core/src/main/java/net/librec/recommender/AbstractRecommender.java:135,
// MatrixRecommender.java:153,260 leave open; selected with rec.recommender.class=<FQCN> through
// util/DriverClassUtil.java:79-88 (a value containing '.' is Class.forName'd) -- zero changes to core.
package net.librec.recommender.cuda;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.DoubleBuffer;
import java.util.ArrayList;

import net.librec.common.LibrecException;
import net.librec.data.structure.AbstractBaseDataEntry;
import net.librec.data.structure.BaseRankingDataEntry;
import net.librec.data.structure.LibrecDataList;
import net.librec.math.structure.DenseMatrix;
import net.librec.math.structure.VectorBasedDenseVector;
import net.librec.recommender.MatrixFactorizationRecommender;
import net.librec.recommender.item.RecommendedList;

abstract class MatrixFactorizationCudaRecommender extends MatrixFactorizationRecommender {
    protected long handle;
    protected VectorBasedDenseVector userBiases, itemBiases;   // only BiasedMF allocates them
    protected double regBias;

    abstract int model();

    @Override
    protected void setup() throws LibrecException {
        super.setup();                                     // MatrixFactorizationRecommender.java:67-94 (Gaussian init on the JVM RNG)
        beforeStage();                                     // WRMF / eALS: weightMatrix() rewrites the train values first
        int device = conf.getInt("rec.cuda.device", 0);
        int mode = "reference".equals(conf.get("rec.cuda.order", "shuffled"))
                ? LibrecB200.UPDATE_REFERENCE_ORDER : LibrecB200.UPDATE_ATOMIC;
        String devs = conf.get("rec.cuda.devices");                    // e.g. 0,1,2,3,4,5,6,7: one handle, all listed GPUs of the box
        if (devs != null && !devs.trim().isEmpty()) {
            String[] parts = devs.trim().split("\\s*,\\s*");
            int[] ids = new int[parts.length];
            for (int d = 0; d < parts.length; d++) ids[d] = Integer.parseInt(parts[d]);
            handle = LibrecB200.createMulti(ids, model(), numFactors, mode, conf.getLong("rec.cuda.seed", 1L), 0);
        } else
        handle = LibrecB200.create(device, model(), numFactors, mode, conf.getLong("rec.cuda.seed", 1L), 0);
        if (handle == 0) throw new LibrecException(LibrecB200.lastError(0));
        // flatten SequentialAccessSparseMatrix (per-row int[]/double[]) into rowptr/col/val once
        ByteBuffer rowptr = LibrecB200.hostAlloc(8L * (numUsers + 1)).order(ByteOrder.nativeOrder());
        ByteBuffer col = LibrecB200.hostAlloc(4L * numRates).order(ByteOrder.nativeOrder());
        ByteBuffer val = LibrecB200.hostAlloc(8L * numRates).order(ByteOrder.nativeOrder());
        long off = 0;
        for (int u = 0; u < numUsers; u++) {
            rowptr.putLong(8 * u, off);
            int[] idx = trainMatrix.row(u).getIndices();           // ascending (VectorBasedSequentialSparseVector.java:74-117)
            for (int p = 0; p < idx.length; p++, off++) {
                col.putInt((int) (4 * off), idx[p]);
                val.putDouble((int) (8 * off), trainMatrix.row(u).getAtPosition(p));
            }
        }
        rowptr.putLong(8 * numUsers, off);
        check(LibrecB200.setTrainCsr(handle, numUsers, numItems, rowptr, col, val));
        LibrecB200.hostFree(rowptr); LibrecB200.hostFree(col); LibrecB200.hostFree(val);
        afterStage();
    }

    /** hooks around the staging of the train matrix (WRMF / eALS: weights before it, item confidences after it) */
    protected void beforeStage() throws LibrecException { }
    protected void afterStage() throws LibrecException { }
    /** runs first in trainModel() (eALS replaces userFactors by a zero matrix, EALSRecommender.java:125) */
    protected void beforeTrainModel() throws LibrecException { }

    /** hooks for models with matrices beyond P / Q / biases (SVD++: impItemFactors) */
    protected void afterSetFactors() throws LibrecException { }
    protected void afterGetFactors() throws LibrecException { }

    /** the reference's own recommendRank (for models the native top-N does not cover) */
    protected RecommendedList recommendRankReference(LibrecDataList<AbstractBaseDataEntry> dataList) throws LibrecException {
        return super.recommendRank(dataList);
    }

    protected void check(int status) throws LibrecException {
        if (status != 0) throw new LibrecException(LibrecB200.lastError(handle));
    }

    private ByteBuffer flatten(DenseMatrix m) {
        ByteBuffer b = LibrecB200.hostAlloc(8L * m.rowSize() * m.columnSize()).order(ByteOrder.nativeOrder());
        DoubleBuffer d = b.asDoubleBuffer();               // JDK 8: Buffer.position(int) returns Buffer, so no call chaining
        for (int r = 0; r < m.rowSize(); r++) { d.position(r * m.columnSize()); d.put(m.getValues()[r]); }
        return b;
    }

    private void unflatten(ByteBuffer b, DenseMatrix m) {
        DoubleBuffer d = b.asDoubleBuffer();
        for (int r = 0; r < m.rowSize(); r++) { d.position(r * m.columnSize()); d.get(m.getValues()[r]); }
    }

    /** trainModel(): the iteration loop, isConverged and updateLRate stay in Java (BiasedMFRecommender.java:101-105). */
    @Override
    protected void trainModel() throws LibrecException {
        beforeTrainModel();
        ByteBuffer P = flatten(userFactors), Q = flatten(itemFactors);
        ByteBuffer bu = userBiases == null ? null : ByteBuffer.allocateDirect(8 * numUsers).order(ByteOrder.nativeOrder());
        ByteBuffer bi = itemBiases == null ? null : ByteBuffer.allocateDirect(8 * numItems).order(ByteOrder.nativeOrder());
        if (bu != null) bu.asDoubleBuffer().put(userBiases.getValues());
        if (bi != null) bi.asDoubleBuffer().put(itemBiases.getValues());                // GBPR has item biases only
        check(LibrecB200.setFactors(handle, P, Q, bu, bi, globalMean));
        afterSetFactors();
        double[] lossOut = new double[1];
        boolean boldDriver = conf.getBoolean("rec.learnrate.bolddriver", false);
        if (!earlyStop && !boldDriver && !verbose) {
            // no decision between iterations (the shipped *-test.properties): one native call for all of them; the decay branch of
            // updateLRate runs natively (include/librec_b200.h, lrk_sgd_epochs), the NaN check of isConverged on every returned loss
            double[] losses = new double[numIterations];
            int st = LibrecB200.sgdEpochs(handle, numIterations, learnRate, decay, maxLearnRate, regUser, regItem, regBias, 1, losses);
            for (int iter = 1; iter <= numIterations; iter++) {
                loss = losses[iter - 1];
                if (Double.isNaN(loss) || Double.isInfinite(loss))
                    throw new LibrecException("Loss = NaN or Infinity: current settings does not fit the recommender! Change the settings and try again!");
                lastLoss = loss;
            }
            if (st != 0 && st != LibrecB200.ERR_DIVERGED) check(st);
        } else
        for (int iter = 1; iter <= numIterations; iter++) {
            int st = LibrecB200.sgdEpoch(handle, learnRate, regUser, regItem, regBias, iter, lossOut);
            loss = lossOut[0];
            if (st != 0 && st != LibrecB200.ERR_DIVERGED) check(st);   // diverged: let isConverged throw the reference's exception
            if (isConverged(iter) && earlyStop) break;     // AbstractRecommender.java:249-267 (throws on NaN/Inf)
            updateLRate(iter);                             // MatrixFactorizationRecommender.java:121-139
        }
        check(LibrecB200.getFactors(handle, P, Q, bu, bi));
        afterGetFactors();
        unflatten(P, userFactors); unflatten(Q, itemFactors);
        if (bu != null) bu.asDoubleBuffer().get(userBiases.getValues());
        if (bi != null) bi.asDoubleBuffer().get(itemBiases.getValues());
        LibrecB200.hostFree(P); LibrecB200.hostFree(Q);
        // inherited predict()/recommendRating()/evaluators keep working on the DenseMatrix copies
    }

    /** recommendRank(): one batched native call replaces the parallelStream loop (MatrixRecommender.java:153-201). */
    @Override
    public RecommendedList recommendRank(LibrecDataList<AbstractBaseDataEntry> dataList) throws LibrecException {
        int n = dataList.size();
        int[] users = new int[n];
        for (int c = 0; c < n; c++) users[c] = ((BaseRankingDataEntry) dataList.getDataEntry(c)).getUserId();
        ByteBuffer items = LibrecB200.hostAlloc(4L * n * topN).order(ByteOrder.nativeOrder());
        ByteBuffer scores = LibrecB200.hostAlloc(8L * n * topN).order(ByteOrder.nativeOrder());
        ByteBuffer counts = LibrecB200.hostAlloc(4L * n).order(ByteOrder.nativeOrder());
        check(LibrecB200.topn(handle, users, n, topN, 1, items, scores, counts));
        RecommendedList list = new RecommendedList(numUsers);
        for (int c = 0; c < n; c++) {
            list.addList(new ArrayList<>());
            int cnt = counts.getInt(4 * c);
            for (int t = 0; t < cnt; t++)                                                             // RecommendedList.java:134-151
                list.add(c, items.getInt(4 * (c * topN + t)), scores.getDouble(8 * (c * topN + t)));
        }
        LibrecB200.hostFree(items); LibrecB200.hostFree(scores); LibrecB200.hostFree(counts);
        if (list.size() == 0) throw new IndexOutOfBoundsException("No item is recommended, there is something error in the recommendation algorithm! Please check it!");
        return list;
    }

    @Override
    protected void cleanup() throws LibrecException {
        LibrecB200.destroy(handle);
        handle = 0;
    }
}
