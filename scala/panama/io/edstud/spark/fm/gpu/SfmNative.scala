package io.edstud.spark.fm.gpu

// ALTERNATIVE binding for a JDK 22+ build (NOT what build.sbt:7-11 pins -- the JNI binding
// scala/io/edstud/spark/fm/gpu/SfmJni.scala + jni/sfm_jni.c is the one the shim uses).
// Panama (java.lang.foreign, JDK 22+) binding of libsparkfm_b200.so -- the reference-side stub a
// maintainer adds; one MethodHandle per C-ABI export used by the Scala layer.
// UNVERIFIED SOURCE: no JVM / scalac exists in the build image (SURVEY.md F3); the same ABI is
// exercised from Python ctypes (sparkfm_b200/_lib.py) and the signatures below mirror
// include/sparkfm_b200.h one to one.

import java.lang.foreign._
import java.lang.foreign.ValueLayout._
import java.lang.invoke.MethodHandle

object SfmNative {
    private val linker = Linker.nativeLinker()
    private val arena  = Arena.global()
    private val lib    = SymbolLookup.libraryLookup(
        sys.props.getOrElse("sparkfm.b200.lib", "libsparkfm_b200.so"), arena)

    private def fn(name: String, res: MemoryLayout, args: MemoryLayout*): MethodHandle =
        linker.downcallHandle(lib.find(name).orElseThrow(), FunctionDescriptor.of(res, args: _*))

    // struct sfm_config (64 bytes; offsets checked by tests/test_host.py)
    val CONFIG: StructLayout = MemoryLayout.structLayout(
        JAVA_INT.withName("abi_version"), JAVA_INT.withName("task"), JAVA_INT.withName("k"),
        JAVA_INT.withName("k0"), JAVA_INT.withName("k1"), JAVA_INT.withName("device"),
        JAVA_LONG.withName("n_slots"), JAVA_FLOAT.withName("reg0"), JAVA_FLOAT.withName("regw"),
        JAVA_FLOAT.withName("regv"), JAVA_FLOAT.withName("step_size"),
        JAVA_FLOAT.withName("mini_batch_fraction"), JAVA_INT.withName("sampler_mode"),
        JAVA_LONG.withName("sampler_seed"))

    val create       = fn("sfm_create", JAVA_INT, ADDRESS, ADDRESS)
    val destroy      = fn("sfm_destroy", JAVA_INT, ADDRESS)
    val lastError    = fn("sfm_last_error", ADDRESS, ADDRESS)
    val setHyper     = fn("sfm_set_hyper", JAVA_INT, ADDRESS, JAVA_FLOAT, JAVA_FLOAT, JAVA_FLOAT,
                          JAVA_FLOAT, JAVA_FLOAT)
    val initModel    = fn("sfm_init_model", JAVA_INT, ADDRESS, JAVA_DOUBLE, JAVA_DOUBLE, JAVA_LONG)
    val setModelF64  = fn("sfm_set_model_f64", JAVA_INT, ADDRESS, JAVA_DOUBLE, ADDRESS, ADDRESS)
    val getModelF64  = fn("sfm_get_model_f64", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS)
    val save         = fn("sfm_save", JAVA_INT, ADDRESS, ADDRESS)
    val load         = fn("sfm_load", JAVA_INT, ADDRESS, JAVA_INT, ADDRESS)
    val predict      = fn("sfm_predict", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS)
    val loadDataset  = fn("sfm_load_dataset", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS,
                          JAVA_LONG, JAVA_LONG)
    val unloadDataset = fn("sfm_unload_dataset", JAVA_INT, ADDRESS)
    val evaluate     = fn("sfm_evaluate", JAVA_INT, ADDRESS, ADDRESS)
    val trainStep    = fn("sfm_train_step", JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS)
    val trainStepCsr = fn("sfm_train_step_csr", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS,
                          JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS)
    val train        = fn("sfm_train", JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS)
    val commUniqueId = fn("sfm_comm_unique_id", JAVA_INT, ADDRESS)
    val commInit     = fn("sfm_comm_init", JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT)
    val alsSweep     = fn("sfm_als_sweep", JAVA_INT, ADDRESS, JAVA_INT, ADDRESS)       // ALS.learn, one sweep
    val commMode     = fn("sfm_comm_mode", JAVA_INT, ADDRESS, ADDRESS)   // 0 none, 1 NCCL, 2 NVLink peer kernel, 3 row-sharded
    val hostAlloc    = fn("sfm_host_alloc", JAVA_INT, ADDRESS, JAVA_LONG)
    val hostFree     = fn("sfm_host_free", JAVA_INT, ADDRESS)

    /** Non-zero status -> exception, like the reference's `throw new Exception(...)`
      * (DataCollection.scala:36). */
    def check(status: Int, handle: MemorySegment = MemorySegment.NULL): Unit = if (status != 0) {
        val msg = if (handle == MemorySegment.NULL) "" else
            lastError.invoke(handle).asInstanceOf[MemorySegment].reinterpret(4096).getString(0)
        throw new Exception(s"sparkfm_b200 error $status: $msg")
    }
}
