/* COMPILE-CHECK STAND-IN for <jni.h> -- NOT the JDK header, never link a library built with it.
 * The build image has no JDK, so tests/test_host.py compiles jni/sfm_jni.c against this file
 * (-fsyntax-only -DSFM_JNI_COMPILE_CHECK_ONLY) to prove that the glue is valid C and names every
 * export with the argument types include/sparkfm_b200.h declares.  It declares only the JNI
 * types and JNIEnv members sfm_jni.c uses; the function-table LAYOUT is not the real one.  A real
 * build uses the JDK's own header:  gcc -shared -fPIC -I$JAVA_HOME/include
 * -I$JAVA_HOME/include/linux -Iinclude jni/sfm_jni.c -L sparkfm_b200 -lsparkfm_b200 -o libsfm_jni.so */
#ifndef SFM_JNI_COMPILE_CHECK_ONLY
#error "jni/compile_check/jni.h is a syntax-check stand-in; build against the JDK's jni.h"
#endif
#ifndef SFM_JNI_STANDIN_H
#define SFM_JNI_STANDIN_H
#include <stdint.h>
typedef int32_t jint;
typedef int64_t jlong;
typedef float jfloat;
typedef double jdouble;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef jint jsize;
typedef void* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jlongArray;
typedef jarray jfloatArray;
typedef jarray jdoubleArray;
typedef jarray jbyteArray;
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
    void* (*GetDirectBufferAddress)(JNIEnv*, jobject);
    jlong (*GetDirectBufferCapacity)(JNIEnv*, jobject);
    const char* (*GetStringUTFChars)(JNIEnv*, jstring, jboolean*);
    void (*ReleaseStringUTFChars)(JNIEnv*, jstring, const char*);
    jstring (*NewStringUTF)(JNIEnv*, const char*);
    jsize (*GetArrayLength)(JNIEnv*, jarray);
    void (*SetIntArrayRegion)(JNIEnv*, jintArray, jsize, jsize, const jint*);
    void (*SetLongArrayRegion)(JNIEnv*, jlongArray, jsize, jsize, const jlong*);
    void (*SetFloatArrayRegion)(JNIEnv*, jfloatArray, jsize, jsize, const jfloat*);
    void (*SetDoubleArrayRegion)(JNIEnv*, jdoubleArray, jsize, jsize, const jdouble*);
    void (*GetByteArrayRegion)(JNIEnv*, jbyteArray, jsize, jsize, jbyte*);
    void (*SetByteArrayRegion)(JNIEnv*, jbyteArray, jsize, jsize, const jbyte*);
    jobject (*NewDirectByteBuffer)(JNIEnv*, void*, jlong);
};
#endif
