package io.edstud.spark.fm.gpu

import java.lang.foreign._
import java.lang.foreign.ValueLayout._
import breeze.linalg.SparseVector
import io.edstud.spark.DataSet
import io.edstud.spark.Task._
import io.edstud.spark.fm.FMModel

/** FMModel (fm/FMModel.scala:9-65) whose parameters live on one B200.  `w0 / w / v` of the
  * super class are refreshed from the device on `sync()`; `predict` of a single vector stays on
  * the JVM (the inherited code), batched scoring and the metrics of Model.scala:13-30 run on the
  * GPU.  UNVERIFIED SOURCE (no JVM in the build image). */
class GpuFMModel(num_attribute: Int, num_factor: Int, task: Task = Regression,
                 init_mean: Double = 0, init_stdev: Double = 0.01, seed: Long = 0,
                 device: Int = 0, samplerSeed: Long = 42)
        extends FMModel(num_attribute, num_factor, init_mean, init_stdev, seed) {

    private val arena = Arena.ofShared()
    val handle: MemorySegment = {
        val cfg = arena.allocate(SfmNative.CONFIG)
        cfg.set(JAVA_INT, 0, 1)                                   // abi_version
        cfg.set(JAVA_INT, 4, if (task == Classification) 1 else 0)
        cfg.set(JAVA_INT, 8, num_factor)
        cfg.set(JAVA_INT, 12, if (k0) 1 else 0)
        cfg.set(JAVA_INT, 16, if (k1) 1 else 0)
        cfg.set(JAVA_INT, 20, device)
        cfg.set(JAVA_LONG, 24, (num_attribute + 1).toLong)        // n_slots (FMModel.scala:18)
        cfg.set(JAVA_FLOAT, 32, reg0.toFloat); cfg.set(JAVA_FLOAT, 36, regw.toFloat)
        cfg.set(JAVA_FLOAT, 40, regv.toFloat); cfg.set(JAVA_FLOAT, 44, 0.1f)
        cfg.set(JAVA_FLOAT, 48, 1.0f); cfg.set(JAVA_LONG, 56, samplerSeed)
        val out = arena.allocate(ADDRESS)
        SfmNative.check(SfmNative.create.invoke(cfg, out).asInstanceOf[Int])
        val h = out.get(ADDRESS, 0)
        SfmNative.check(SfmNative.initModel.invoke(h, init_mean, init_stdev, seed).asInstanceOf[Int], h)
        h
    }

    private var resident: DataSet = null

    /** DataSet.cache() (DataSet.scala:50-54): pack the RDD rows into CSR once and copy them to
      * the device.  Packing keeps stored order and duplicates (activeIterator semantics). */
    def cache(dataset: DataSet): Unit = if (resident ne dataset) {
        val rows = dataset.rdd.collect()                           // driver-side pack
        val n = rows.length
        val nnz = rows.map(_._2.activeSize.toLong).sum
        val a = Arena.ofConfined()
        try {
            val rowPtr = a.allocate(JAVA_LONG, n + 1L); val idx = a.allocate(JAVA_INT, math.max(nnz, 1L))
            val value = a.allocate(JAVA_FLOAT, math.max(nnz, 1L)); val label = a.allocate(JAVA_FLOAT, math.max(n, 1).toLong)
            var p = 0L
            for (r <- 0 until n) {
                rowPtr.setAtIndex(JAVA_LONG, r.toLong, p)
                label.setAtIndex(JAVA_FLOAT, r.toLong, rows(r)._1.toFloat)
                rows(r)._2.activeIterator.foreach { case (i, x) =>
                    idx.setAtIndex(JAVA_INT, p, i); value.setAtIndex(JAVA_FLOAT, p, x.toFloat); p += 1 }
            }
            rowPtr.setAtIndex(JAVA_LONG, n.toLong, p)
            SfmNative.check(SfmNative.loadDataset.invoke(handle, rowPtr, idx, value, label, n.toLong, 0L)
                .asInstanceOf[Int], handle)
            resident = dataset
        } finally a.close()
    }

    /** Model.computeRMSE (Model.scala:13-19) as one fused device pass. */
    override def computeRMSE(dataset: DataSet): Double = {
        cache(dataset)
        val a = Arena.ofConfined()
        try {
            val m = a.allocate(JAVA_DOUBLE, 5L)
            SfmNative.check(SfmNative.evaluate.invoke(handle, m).asInstanceOf[Int], handle)
            val rmse = m.getAtIndex(JAVA_DOUBLE, 0L)
            logInfo(dataset.rdd.name + " RMSE = " + rmse)
            rmse
        } finally a.close()
    }

    /** Refresh the JVM copies of w0 / w / v (fm/FMModel.scala:17-19) from the device. */
    def sync(): Unit = {
        val a = Arena.ofConfined()
        try {
            val n = num_attribute + 1
            val w0s = a.allocate(JAVA_DOUBLE); val ws = a.allocate(JAVA_DOUBLE, n.toLong)
            val vs = a.allocate(JAVA_DOUBLE, n.toLong * num_factor)
            SfmNative.check(SfmNative.getModelF64.invoke(handle, w0s, ws, vs).asInstanceOf[Int], handle)
            w0 = w0s.get(JAVA_DOUBLE, 0)
            for (i <- 0 until n) { w(i) = ws.getAtIndex(JAVA_DOUBLE, i.toLong)
                for (f <- 0 until num_factor) v(f, i) = vs.getAtIndex(JAVA_DOUBLE, i.toLong * num_factor + f) }
        } finally a.close()
    }

    def save(path: String): Unit = {
        val a = Arena.ofConfined()
        try SfmNative.check(SfmNative.save.invoke(handle, a.allocateFrom(path)).asInstanceOf[Int], handle)
        finally a.close()
    }

    def close(): Unit = { SfmNative.destroy.invoke(handle); arena.close() }
}
