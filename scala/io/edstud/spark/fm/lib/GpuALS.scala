package io.edstud.spark.fm.lib

import io.edstud.spark.DataSet
import io.edstud.spark.fm._
import io.edstud.spark.fm.gpu._

/** The reference's own learner (fm/lib/ALS.scala:11-208) with the sweep on the GPU: same plugin
  * boundary, same coordinate order, same closed-form update, the residual map `e` and the
  * per-factor map `q` kept on the device instead of in driver hash maps.  Used like the
  * original: `FM(dataset, k, task, iters).learnWith(GpuALS.run())`.  `refQuirks = true` skips the
  * last slot like `0 until num_attribute` does (ALS.scala:38,52; see sfm_als_sweep in
  * include/sparkfm_b200.h).  UNVERIFIED SOURCE (no JVM in the build image). */
class GpuALS protected (val refQuirks: Boolean) extends FMLearn {

    val rmseHistory = scala.collection.mutable.ArrayBuffer[Double]()

    override def learn(fm: FMModel, dataset: DataSet): FMModel = {
        val gpu = fm match {
            case g: GpuFMModel => g
            case _ => throw new Exception("GpuALS needs a GpuFMModel")
        }
        gpu.cache(dataset)
        // the model's own regularisation (FMModel.scala:29-31), as ALS.scala:21,40,56 reads it
        SfmJni.check(SfmJni.setHyper(gpu.handle, fm.reg0.toFloat, fm.regw.toFloat, fm.regv.toFloat,
            0.1f, 1.0f), gpu.handle)
        val rmse = new Array[Double](1)
        SfmJni.check(SfmJni.alsSweep(gpu.handle, if (refQuirks) 1 else 0, rmse), gpu.handle)
        gpu.markStale()
        rmseHistory += rmse(0)
        logInfo("Finish")
        gpu
    }
}

object GpuALS {
    def run(refQuirks: Boolean = false): GpuALS = new GpuALS(refQuirks)
}
