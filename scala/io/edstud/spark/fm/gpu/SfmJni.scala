package io.edstud.spark.fm.gpu

import java.nio.{Buffer, ByteBuffer, ByteOrder}

/** JNI binding of libsparkfm_b200.so for the JVM the reference pins (Scala 2.10 / Spark 1.2.0,
  * Java 7/8 -- build.sbt:7-11): one `@native` method per C-ABI export of include/sparkfm_b200.h,
  * implemented by jni/sfm_jni.c (libsfm_jni.so, linked against libsparkfm_b200.so).
  *
  * Bulk arrays are direct java.nio buffers in native byte order (null = NULL); scalar outputs are
  * small primitive arrays; every method returns the library's int32 status (0 = ok), `check`
  * turns a non-zero status into the reference's own error idiom, `throw new Exception(...)`
  * (DataCollection.scala:36).  UNVERIFIED SOURCE: no JVM / scalac exists in the build image; the
  * C side is syntax-checked and export-checked by tests/test_host.py, the same ABI runs end to
  * end from Python ctypes. */
object SfmJni {
    System.load(sys.props.getOrElse("sparkfm.b200.jni", "libsfm_jni.so"))

    @native def abiVersion(): Int
    @native def statusString(status: Int): String
    @native def deviceCount(): Int
    @native def hostAlloc(bytes: Long, addrOut: Array[Long]): Int
    @native def wrapHost(addr: Long, bytes: Long): ByteBuffer
    @native def hostFree(addr: Long): Int

    @native def create(task: Int, k: Int, k0: Int, k1: Int, device: Int, nSlots: Long, reg0: Float,
                       regw: Float, regv: Float, stepSize: Float, miniBatchFraction: Float,
                       samplerMode: Int, samplerSeed: Long, handleOut: Array[Long]): Int
    @native def destroy(h: Long): Int
    @native def lastError(h: Long): String
    @native def getConfig(h: Long, ints6: Array[Int], longs2: Array[Long], floats5: Array[Float]): Int
    @native def setHyper(h: Long, reg0: Float, regw: Float, regv: Float, stepSize: Float,
                         miniBatchFraction: Float): Int

    @native def initModel(h: Long, mean: Double, stdev: Double, seed: Long): Int
    @native def setModel(h: Long, w0: Float, w: Buffer, v: Buffer): Int
    @native def getModel(h: Long, w0Out: Array[Float], w: Buffer, v: Buffer): Int
    @native def setModelF64(h: Long, w0: Double, w: Buffer, v: Buffer): Int
    @native def getModelF64(h: Long, w0Out: Array[Double], w: Buffer, v: Buffer): Int
    @native def save(h: Long, path: String): Int
    @native def load(path: String, device: Int, handleOut: Array[Long]): Int

    @native def predict(h: Long, rowPtr: Buffer, idx: Buffer, value: Buffer, nRows: Long, out: Buffer): Int
    @native def loadDataset(h: Long, rowPtr: Buffer, idx: Buffer, value: Buffer, label: Buffer,
                            nRows: Long, globalRowOffset: Long): Int
    @native def unloadDataset(h: Long): Int
    @native def synthCtrDataset(h: Long, nRows: Long, globalRowOffset: Long, nFields: Int,
                                fieldLog2Card: Buffer, zipfCdf: Buffer, zipfCdfOff: Buffer, seed: Long): Int
    @native def getDatasetRows(h: Long, rowLo: Long, rowHi: Long, rowPtr: Buffer, idx: Buffer,
                               value: Buffer, label: Buffer): Int
    @native def datasetInfo(h: Long, rowsNnzOut: Array[Long], maxIndexOut: Array[Int]): Int
    @native def predictResident(h: Long, rowLo: Long, rowHi: Long, out: Buffer): Int
    @native def evaluate(h: Long, metrics5: Array[Double]): Int
    @native def evaluateAuc(h: Long, out3: Array[Double]): Int

    @native def trainStep(h: Long, rowIds: Buffer, nIds: Long, iter: Long, lossOut: Array[Double],
                          batchOut: Array[Long]): Int
    @native def trainStepCsr(h: Long, rowPtr: Buffer, idx: Buffer, value: Buffer, label: Buffer,
                             nRows: Long, iter: Long, lossOut: Array[Double], batchOut: Array[Long]): Int
    @native def stageCsr(h: Long, slot: Int, rowPtr: Buffer, idx: Buffer, value: Buffer, label: Buffer,
                         nRows: Long): Int
    @native def trainStepStaged(h: Long, slot: Int, iter: Long, lossOut: Array[Double],
                                batchOut: Array[Long]): Int
    @native def stageOnehot(h: Long, slot: Int, packedIdx: Buffer, labelBits: Buffer, labelF32: Buffer,
                            nRows: Long, m: Int, idBits: Int): Int
    @native def packOnehot(idx: Buffer, label: Buffer, nRows: Long, m: Int, idBits: Int,
                           packedIdx: Buffer, labelBits: Buffer): Int
    @native def train(h: Long, firstIter: Long, nIters: Long, lossHistory: Buffer): Int
    @native def sampleRows(seed: Long, iter: Long, fraction: Double, rowLo: Long, rowHi: Long,
                           out: Buffer, nOut: Array[Long]): Int
    @native def partitionRows(seed: Long, nParts: Long, part: Long, rowLo: Long, rowHi: Long,
                              out: Buffer, nOut: Array[Long]): Int
    @native def gradient(h: Long, rowIds: Buffer, nIds: Long, gradV: Buffer, gradW: Buffer,
                         gradW0Out: Array[Float], lossSumOut: Array[Double], batchOut: Array[Long]): Int

    @native def alsSweep(h: Long, flags: Int, rmseOut: Array[Double]): Int
    @native def alsResiduals(h: Long, out: Buffer, n: Long): Int

    @native def commUniqueId(id128: Array[Byte]): Int
    @native def commInit(h: Long, id128: Array[Byte], rank: Int, worldSize: Int): Int
    @native def commInfo(h: Long, rankWorldOut: Array[Int]): Int
    @native def commMode(h: Long, modeOut: Array[Int]): Int
    @native def commBroadcastModel(h: Long): Int

    @native def parseLibfm(text: Buffer, len: Long, numFeatures: Int, counts4Out: Array[Long],
                           label: Buffer, rowPtr: Buffer, idx: Buffer, value: Buffer): Int
    @native def formatLibfm(label: Buffer, rowPtr: Buffer, idx: Buffer, value: Buffer, nRows: Long,
                            out: Buffer, cap: Long, neededOut: Array[Long]): Int

    @native def statsGet(h: Long, counters8: Array[Long], ms7: Array[Double]): Int
    @native def statsReset(h: Long): Int
    @native def setPhaseTiming(h: Long, enabled: Int): Int
    @native def synchronize(h: Long): Int
    @native def timerStart(h: Long): Int
    @native def timerStop(h: Long, msOut: Array[Float]): Int

    /** Direct buffer in native byte order (what the C side reads through GetDirectBufferAddress). */
    def direct(bytes: Long): ByteBuffer =
        ByteBuffer.allocateDirect(math.max(bytes, 8L).toInt).order(ByteOrder.nativeOrder())

    def check(status: Int, h: Long = 0L): Unit = if (status != 0) {
        val msg = if (h == 0L) statusString(status) else statusString(status) + ": " + lastError(h)
        throw new Exception("sparkfm_b200 error " + status + ": " + msg)
    }
}
