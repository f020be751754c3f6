package io.edstud.spark.fm.gpu

import breeze.linalg.SparseVector
import io.edstud.spark.DataSet
import io.edstud.spark.Task._
import io.edstud.spark.fm.FMModel

/** FMModel (fm/FMModel.scala:9-65) whose parameters live on one B200, bound through JNI
  * (SfmJni.scala; Scala 2.10 / Java 7-8 as build.sbt:7-11 pins).
  *
  * The device is the source of truth.  The JVM-side `w0 / w / v` of the super class are the
  * device's values whenever a caller can observe them: the constructor pushes the JVM Gaussian
  * draw (FMModel.scala:19-22) to the device, so both sides start from the same model, and every
  * learner (`SGD.learn`, `GpuALS.learn`) marks the JVM copy stale; `predict`, `computeMAE` and
  * `computeAccuracy` (Model.scala:11-30) refresh it lazily before the inherited code runs, and
  * `computeRMSE` runs on the device.  UNVERIFIED SOURCE (no JVM in the build image). */
class GpuFMModel(num_attribute: Int, num_factor: Int, task: Task = Regression,
                 init_mean: Double = 0, init_stdev: Double = 0.01, seed: Long = 0,
                 device: Int = 0, samplerSeed: Long = 42)
        extends FMModel(num_attribute, num_factor, init_mean, init_stdev, seed) {

    private val nSlots = num_attribute + 1                          // FMModel.scala:18

    val handle: Long = {
        val out = new Array[Long](1)
        SfmJni.check(SfmJni.create(if (task == Classification) 1 else 0, num_factor,
            if (k0) 1 else 0, if (k1) 1 else 0, device, nSlots.toLong, reg0.toFloat, regw.toFloat,
            regv.toFloat, 0.1f, 1.0f, 0, samplerSeed, out))
        out(0)
    }
    push()                                                          // same initial model on both sides

    @transient private var stale = false
    @transient private var resident: DataSet = null

    /** A learner changed the device model: the JVM fields are refreshed at their next use. */
    def markStale(): Unit = stale = true

    /** JVM w0 / w / v (Breeze DenseMatrix(k, n+1), column-major = feature-major [n+1][k]) -> device. */
    def push(): Unit = {
        val wb = SfmJni.direct(8L * nSlots); val vb = SfmJni.direct(8L * nSlots * num_factor)
        val wd = wb.asDoubleBuffer(); val vd = vb.asDoubleBuffer()
        for (i <- 0 until nSlots) { wd.put(i, w(i)); for (f <- 0 until num_factor) vd.put(i * num_factor + f, v(f, i)) }
        SfmJni.check(SfmJni.setModelF64(handle, w0, wb, vb), handle)
        stale = false
    }

    /** Refresh the JVM copies of w0 / w / v (fm/FMModel.scala:17-19) from the device. */
    def sync(): Unit = {
        val wb = SfmJni.direct(8L * nSlots); val vb = SfmJni.direct(8L * nSlots * num_factor)
        val w0o = new Array[Double](1)
        SfmJni.check(SfmJni.getModelF64(handle, w0o, wb, vb), handle)
        val wd = wb.asDoubleBuffer(); val vd = vb.asDoubleBuffer()
        w0 = w0o(0)
        for (i <- 0 until nSlots) { w(i) = wd.get(i); for (f <- 0 until num_factor) v(f, i) = vd.get(i * num_factor + f) }
        stale = false
    }

    /** DataSet.cache() (DataSet.scala:50-54): pack the RDD rows into CSR once and copy them to
      * the device.  Packing keeps stored order and duplicates (activeIterator semantics). */
    def cache(dataset: DataSet): Unit = if (resident ne dataset) {
        val rows = dataset.rdd.collect()                            // driver-side pack
        val n = rows.length
        val nnz = rows.map(_._2.activeSize.toLong).sum
        val rowPtr = SfmJni.direct(8L * (n + 1)); val idx = SfmJni.direct(4L * nnz)
        val value = SfmJni.direct(4L * nnz); val label = SfmJni.direct(4L * n)
        val rp = rowPtr.asLongBuffer(); val ix = idx.asIntBuffer()
        val vx = value.asFloatBuffer(); val lb = label.asFloatBuffer()
        var p = 0
        for (r <- 0 until n) {
            rp.put(r, p.toLong)
            lb.put(r, rows(r)._1.toFloat)
            rows(r)._2.activeIterator.foreach { case (i, x) => ix.put(p, i); vx.put(p, x.toFloat); p += 1 }
        }
        rp.put(n, p.toLong)
        SfmJni.check(SfmJni.loadDataset(handle, rowPtr, idx, value, label, n.toLong, 0L), handle)
        resident = dataset
    }

    /** Model.predict (Model.scala:11) of one vector: the inherited JVM code on fresh parameters. */
    override def predict(features: SparseVector[Double]): Double = {
        if (stale) sync()
        super.predict(features)
    }

    /** Model.computeRMSE (Model.scala:13-19) as one fused device pass. */
    override def computeRMSE(dataset: DataSet): Double = {
        cache(dataset)
        val m = new Array[Double](5)                                // rmse, "mae", accuracy, logloss, n
        SfmJni.check(SfmJni.evaluate(handle, m), handle)
        logInfo(dataset.rdd.name + " RMSE = " + m(0))
        m(0)
    }

    /** Model.computeMAE / computeAccuracy (Model.scala:21-30): the reference's own RDD code, its
      * quirks included (no abs, integer division), over up-to-date parameters. */
    override def computeMAE(dataset: DataSet): Double = { if (stale) sync(); super.computeMAE(dataset) }
    override def computeAccuracy(dataset: DataSet): Double = { if (stale) sync(); super.computeAccuracy(dataset) }

    def save(path: String): Unit = SfmJni.check(SfmJni.save(handle, path), handle)

    def close(): Unit = SfmJni.destroy(handle)
}
