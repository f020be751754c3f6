package io.edstud.spark.fm.lib

import io.edstud.spark.DataSet
import io.edstud.spark.fm._
import io.edstud.spark.fm.gpu._

/** Mini-batch SGD learner behind the reference's plugin boundary (fm/FMLearn.scala:10-16); used
  * exactly like ALS: `FM(dataset, k, task, iters).learnWith(SGD.run(...))`.  One `learn` call is
  * one iteration: Bernoulli(miniBatchFraction) batch, gradient sum, eta = stepSize/sqrt(t), L2
  * regParam = (r0, r1, r2) -- all on the GPU (DESIGN.md section 2).  UNVERIFIED SOURCE. */
class SGD protected (val stepSize: Double, val regParam: (Double, Double, Double),
                     val miniBatchFraction: Double) extends FMLearn {

    private var iteration = 0L
    val lossHistory = scala.collection.mutable.ArrayBuffer[Double]()

    override def learn(fm: FMModel, dataset: DataSet): FMModel = {
        val gpu = fm match {
            case g: GpuFMModel => g
            case _ => throw new Exception("SGD needs a GpuFMModel (build it with FMWithSGD or GpuFM)")
        }
        gpu.cache(dataset)
        SfmJni.check(SfmJni.setHyper(gpu.handle, regParam._1.toFloat, regParam._2.toFloat,
            regParam._3.toFloat, stepSize.toFloat, miniBatchFraction.toFloat), gpu.handle)
        iteration += 1
        val loss = new Array[Double](1); val batch = new Array[Long](1)
        // row_ids = NULL, n_ids = -1: the built-in sampler
        SfmJni.check(SfmJni.trainStep(gpu.handle, null, -1L, iteration, loss, batch), gpu.handle)
        gpu.markStale()                                   // predict / computeMAE re-read the device model
        lossHistory += loss(0)
        logDebug("SGD iteration " + iteration + ": mean loss " + lossHistory.last)
        gpu
    }
}

object SGD {
    def run(stepSize: Double = 0.1, regParam: (Double, Double, Double) = (0.0, 0.0, 0.0),
            miniBatchFraction: Double = 1.0): SGD = new SGD(stepSize, regParam, miniBatchFraction)
}
