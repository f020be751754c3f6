package io.edstud.spark.fm

import org.apache.spark.rdd.RDD
import org.apache.spark.mllib.regression.LabeledPoint
import org.apache.spark.mllib.linalg.{SparseVector => MLSparse}
import breeze.linalg.SparseVector
import io.edstud.spark.DataSet
import io.edstud.spark.Task._
import io.edstud.spark.fm.gpu.GpuFMModel
import io.edstud.spark.fm.lib.SGD

/** spark-libFM style entry point named by BASELINE.json north_star (not in the reference):
  * RDD[LabeledPoint] in, trained FMModel out.  dim = (k0, k1, k), regParam = (r0, r1, r2).
  * UNVERIFIED SOURCE. */
object FMWithSGD {
    def train(input: RDD[LabeledPoint], task: Task, numIterations: Int, stepSize: Double,
              miniBatchFraction: Double, dim: (Boolean, Boolean, Int),
              regParam: (Double, Double, Double), initStd: Double): FMModel = {
        val rows = input.map { lp =>
            val sv = lp.features.asInstanceOf[MLSparse]
            (lp.label, new SparseVector[Double](sv.indices, sv.values, sv.size))
        }
        val dataset = DataSet("FMWithSGD.input", rows).cache()
        var fm: FMModel = new GpuFMModel(dataset.dimension, dim._3, task, 0.0, initStd)
        val learner = SGD.run(stepSize, regParam, miniBatchFraction)
        for (i <- 1 to numIterations) fm = learner.learn(fm, dataset)   // FactorizationMachines.scala:42-46
        fm.asInstanceOf[GpuFMModel].sync()
        dataset.unpersist()
        fm
    }
}
