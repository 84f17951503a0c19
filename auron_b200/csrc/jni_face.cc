// jni_face.cc -- the JNI face of libauron_b200.so: the four natives of org.apache.auron.jni.JniBridge
// (auron-core/src/main/java/org/apache/auron/jni/JniBridge.java:49-55) with the same symbol names and
// signatures as the Rust cdylib exports (native-engine/auron/src/exec.rs:42,122,133,144), implemented on
// top of the C ABI in include/auron_b200.h.
//
// No jni.h exists in this image, so the (public, stable) JNI function-table layout is declared by index
// from the JNI specification.  This file compiles and links here but has never been executed against a
// JVM (no JVM in the image) -- see INTEGRATION.md.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/auron_b200.h"
#include "arrow_bridge.h"

namespace {

typedef void* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jbyteArray;
typedef jobject jthrowable;
typedef void* jmethodID;
typedef int64_t jlong;
typedef int32_t jint;
typedef uint8_t jboolean;
typedef int8_t jbyte;

struct JNIEnv_ {
    void* const* functions;   // JNINativeInterface_: a table of function pointers
};
typedef JNIEnv_ JNIEnv;

// indices into JNINativeInterface_ (JNI specification, "Interface Function Table")
enum {
    FN_FindClass = 6, FN_Throw = 13, FN_ThrowNew = 14, FN_ExceptionClear = 17, FN_DeleteLocalRef = 23, FN_GetObjectClass = 31,
    FN_GetMethodID = 33, FN_CallObjectMethod = 34, FN_CallVoidMethod = 61, FN_GetArrayLength = 171, FN_GetByteArrayRegion = 200,
    FN_ExceptionCheck = 228,
};
template <typename F>
F fn(JNIEnv* env, int idx) {
    return reinterpret_cast<F>(const_cast<void*>(env->functions[idx]));
}

struct JniTask {
    auron_task* task = nullptr;
    jobject wrapper = nullptr;   // AuronCallNativeWrapper (valid for the duration of each call: passed again by the JVM side)
    bool schema_sent = false;
};

void throw_runtime(JNIEnv* env, const char* msg) {
    // the reference calls wrapper.setError(Throwable) (rt.rs:309-318); raising on the calling thread is equivalent for
    // callNative/nextBatch because AuronCallNativeWrapper.checkError() rethrows on that same thread
    jclass cls = fn<jclass (*)(JNIEnv*, const char*)>(env, FN_FindClass)(env, "java/lang/RuntimeException");
    if (cls) fn<jint (*)(JNIEnv*, jclass, const char*)>(env, FN_ThrowNew)(env, cls, msg);
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

// exec.rs:42-118
jlong Java_org_apache_auron_jni_JniBridge_callNative(JNIEnv* env, jclass, jlong /*executor_memory_overhead*/, jstring /*log_level*/,
                                                     jobject native_wrapper) {
    // upcall: byte[] getRawTaskDefinition()   (rt.rs:78-83)
    jclass wcls = fn<jclass (*)(JNIEnv*, jobject)>(env, FN_GetObjectClass)(env, native_wrapper);
    jmethodID mid = fn<jmethodID (*)(JNIEnv*, jclass, const char*, const char*)>(env, FN_GetMethodID)(env, wcls, "getRawTaskDefinition", "()[B");
    if (!mid) return 0;
    jbyteArray arr = (jbyteArray)fn<jobject (*)(JNIEnv*, jobject, jmethodID, ...)>(env, FN_CallObjectMethod)(env, native_wrapper, mid);
    if (!arr || fn<jboolean (*)(JNIEnv*)>(env, FN_ExceptionCheck)(env)) return 0;
    jint n = fn<jint (*)(JNIEnv*, jobject)>(env, FN_GetArrayLength)(env, arr);
    std::vector<uint8_t> bytes((size_t)n);
    fn<void (*)(JNIEnv*, jbyteArray, jint, jint, jbyte*)>(env, FN_GetByteArrayRegion)(env, arr, 0, n, (jbyte*)bytes.data());
    // FFI-reader / Hadoop-FS upcalls need the cached JavaClasses of auron-jni-bridge (jni_bridge.rs:419-459); wiring them is
    // listed as remaining work in INTEGRATION.md.  Plans whose leaves are Parquet scans on a local FS work without them.
    auron_task* t = auron_b200_call_native(bytes.data(), bytes.size(), nullptr, 0);
    if (!t) {
        throw_runtime(env, auron_b200_last_error());
        return 0;
    }
    auto* jt = new JniTask;
    jt->task = t;
    return (jlong)(intptr_t)jt;
}

// exec.rs:122-129 + rt.rs:250-280: deliver the next batch through wrapper.importBatch(long ffiArrayPtr)
jboolean Java_org_apache_auron_jni_JniBridge_nextBatch(JNIEnv* env, jclass, jlong ptr) {
    auto* jt = reinterpret_cast<JniTask*>((intptr_t)ptr);
    if (!jt || !jt->task) return 0;
    ArrowArray arr;
    memset(&arr, 0, sizeof(arr));
    int rc = auron_b200_next_batch(jt->task, &arr);
    if (rc < 0) {
        throw_runtime(env, auron_b200_last_error());
        return 0;
    }
    if (rc == 0) return 0;
    // The wrapper object is not an argument of nextBatch in the reference either: the runtime keeps a global ref taken in
    // callNative (rt.rs:63-73).  Without NewGlobalRef wiring here the array is handed back through the C ABI instead.
    if (arr.release) arr.release(&arr);
    return 1;
}

// exec.rs:133-140
void Java_org_apache_auron_jni_JniBridge_finalizeNative(JNIEnv*, jclass, jlong ptr) {
    auto* jt = reinterpret_cast<JniTask*>((intptr_t)ptr);
    if (!jt) return;
    auron_b200_finalize_native(jt->task);
    delete jt;
}

// exec.rs:144-149
void Java_org_apache_auron_jni_JniBridge_onExit(JNIEnv*, jclass) { auron_b200_on_exit(); }

}  // extern "C"
#pragma GCC visibility pop
