// jni_face.cc -- the JNI face of libauron_b200.so: the four natives of org.apache.auron.jni.JniBridge
// (auron-core/src/main/java/org/apache/auron/jni/JniBridge.java:49-55) with the same symbol names and
// signatures as the Rust cdylib exports (native-engine/auron/src/exec.rs:42,122,133,144), implemented on
// top of the C ABI in include/auron_b200.h, plus the upcalls the reference makes on this path:
//
//   AuronCallNativeWrapper.getRawTaskDefinition / importSchema / importBatch / setError / getMetrics
//                                          (rt.rs:78-83,167-170,258-262,309-318,300-306; jni_bridge.rs:1485-1525)
//   JniBridge.getResource / isTaskRunning / openFileAsDataInputWrapper / stringConf / intConf /
//             get+setContextClassLoader / get+setThreadContext                (jni_bridge.rs:651-777, rt.rs:117-134)
//   AuronArrowFFIExporter.exportNextBatch + AutoCloseable.close            (ffi_reader_exec.rs:126-130,195,214-217)
//   scala.Function1.apply (fs provider) + FSDataInputWrapper.readFully     (hadoop_fs.rs:55-66,85-96,145-153)
//   scala.Function0.apply -> scala.collection.Iterator of BlockObject      (ipc_reader_exec.rs:147-154,186-207,279-330)
//   MetricNode.getChild / add                                              (metrics.rs:22-58)
//
// No jni.h exists in this image, so the (public, stable) JNI function-table layout is declared by index from the JNI
// specification ("Interface Function Table").  No JVM exists here either: the file is exercised end to end against a
// mock function table (tests/jni_mock/mock_jvm.cc) that implements exactly the classes and methods named above.
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
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
union jvalue {
    jboolean z;
    jbyte b;
    uint16_t c;
    int16_t s;
    jint i;
    jlong j;
    float f;
    double d;
    jobject l;
};

struct JNIEnv_ {
    void* const* functions;   // JNINativeInterface_: a table of function pointers
};
typedef JNIEnv_ JNIEnv;
struct JavaVM_ {
    void* const* functions;   // JNIInvokeInterface_
};
typedef JavaVM_ JavaVM;

// indices into JNINativeInterface_ / JNIInvokeInterface_ (JNI specification)
enum {
    FN_FindClass = 6, FN_Throw = 13, FN_ThrowNew = 14, FN_ExceptionOccurred = 15, FN_ExceptionClear = 17, FN_PushLocalFrame = 19,
    FN_PopLocalFrame = 20, FN_NewGlobalRef = 21, FN_DeleteGlobalRef = 22, FN_NewObjectA = 30, FN_GetObjectClass = 31, FN_GetMethodID = 33,
    FN_CallObjectMethodA = 36, FN_CallBooleanMethodA = 39, FN_CallIntMethodA = 51, FN_CallLongMethodA = 54, FN_CallVoidMethodA = 63,
    FN_GetStaticMethodID = 113, FN_CallStaticObjectMethodA = 116, FN_CallStaticBooleanMethodA = 119, FN_CallStaticIntMethodA = 131, FN_CallStaticVoidMethodA = 143, FN_NewStringUTF = 167,
    FN_GetStringUTFChars = 169, FN_ReleaseStringUTFChars = 170, FN_GetArrayLength = 171, FN_GetByteArrayRegion = 200, FN_GetJavaVM = 219,
    FN_ExceptionCheck = 228, FN_NewDirectByteBuffer = 229, FN_GetDirectBufferAddress = 230,
};
enum { VM_DetachCurrentThread = 5, VM_GetEnv = 6, VM_AttachCurrentThreadAsDaemon = 7 };
constexpr jint kJniVersion = 0x00010008;

struct JavaError {};   // a Java exception is pending on (or was captured from) the current thread

// Thin typed view of one thread's JNIEnv.  Every call that can raise checks ExceptionCheck and throws JavaError.
struct J {
    JNIEnv* env;
    template <typename F>
    F fn(int idx) const {
        return reinterpret_cast<F>(const_cast<void*>(env->functions[idx]));
    }
    bool pending() const { return fn<jboolean (*)(JNIEnv*)>(FN_ExceptionCheck)(env) != 0; }
    void check() const {
        if (pending()) throw JavaError();
    }
    jclass find_class(const char* name) const {
        jclass c = fn<jclass (*)(JNIEnv*, const char*)>(FN_FindClass)(env, name);
        check();
        return c;
    }
    jclass class_of(jobject o) const { return fn<jclass (*)(JNIEnv*, jobject)>(FN_GetObjectClass)(env, o); }
    jmethodID method(jclass c, const char* name, const char* sig) const {
        jmethodID m = fn<jmethodID (*)(JNIEnv*, jclass, const char*, const char*)>(FN_GetMethodID)(env, c, name, sig);
        check();
        return m;
    }
    jmethodID static_method(jclass c, const char* name, const char* sig) const {
        jmethodID m = fn<jmethodID (*)(JNIEnv*, jclass, const char*, const char*)>(FN_GetStaticMethodID)(env, c, name, sig);
        check();
        return m;
    }
    // instance calls resolve the method on the object's own class: same virtual dispatch as the interface lookup of the reference
    jobject call_object(jobject o, const char* name, const char* sig, const jvalue* a = nullptr) const {
        jmethodID m = method(class_of(o), name, sig);
        jobject r = fn<jobject (*)(JNIEnv*, jobject, jmethodID, const jvalue*)>(FN_CallObjectMethodA)(env, o, m, a);
        check();
        return r;
    }
    bool call_bool(jobject o, const char* name, const char* sig, const jvalue* a = nullptr) const {
        jmethodID m = method(class_of(o), name, sig);
        jboolean r = fn<jboolean (*)(JNIEnv*, jobject, jmethodID, const jvalue*)>(FN_CallBooleanMethodA)(env, o, m, a);
        check();
        return r != 0;
    }
    jint call_int(jobject o, const char* name, const char* sig, const jvalue* a = nullptr) const {
        jmethodID m = method(class_of(o), name, sig);
        jint r = fn<jint (*)(JNIEnv*, jobject, jmethodID, const jvalue*)>(FN_CallIntMethodA)(env, o, m, a);
        check();
        return r;
    }
    jlong call_long(jobject o, const char* name, const char* sig, const jvalue* a = nullptr) const {
        jmethodID m = method(class_of(o), name, sig);
        jlong r = fn<jlong (*)(JNIEnv*, jobject, jmethodID, const jvalue*)>(FN_CallLongMethodA)(env, o, m, a);
        check();
        return r;
    }
    void call_void(jobject o, const char* name, const char* sig, const jvalue* a = nullptr) const {
        jmethodID m = method(class_of(o), name, sig);
        fn<void (*)(JNIEnv*, jobject, jmethodID, const jvalue*)>(FN_CallVoidMethodA)(env, o, m, a);
        check();
    }
    jobject call_static_object(jclass c, const char* name, const char* sig, const jvalue* a = nullptr) const {
        jmethodID m = static_method(c, name, sig);
        jobject r = fn<jobject (*)(JNIEnv*, jclass, jmethodID, const jvalue*)>(FN_CallStaticObjectMethodA)(env, c, m, a);
        check();
        return r;
    }
    bool call_static_bool(jclass c, const char* name, const char* sig, const jvalue* a = nullptr) const {
        jmethodID m = static_method(c, name, sig);
        jboolean r = fn<jboolean (*)(JNIEnv*, jclass, jmethodID, const jvalue*)>(FN_CallStaticBooleanMethodA)(env, c, m, a);
        check();
        return r != 0;
    }
    void call_static_void(jclass c, const char* name, const char* sig, const jvalue* a = nullptr) const {
        jmethodID m = static_method(c, name, sig);
        fn<void (*)(JNIEnv*, jclass, jmethodID, const jvalue*)>(FN_CallStaticVoidMethodA)(env, c, m, a);
        check();
    }
    jint call_static_int(jclass c, const char* name, const char* sig, const jvalue* a = nullptr) const {
        jmethodID m = static_method(c, name, sig);
        jint r = fn<jint (*)(JNIEnv*, jclass, jmethodID, const jvalue*)>(FN_CallStaticIntMethodA)(env, c, m, a);
        check();
        return r;
    }
    jobject new_object(jclass c, const char* ctor_sig, const jvalue* a) const {
        jmethodID m = method(c, "<init>", ctor_sig);
        jobject r = fn<jobject (*)(JNIEnv*, jclass, jmethodID, const jvalue*)>(FN_NewObjectA)(env, c, m, a);
        check();
        return r;
    }
    jstring new_string(const char* s) const {
        jstring r = fn<jstring (*)(JNIEnv*, const char*)>(FN_NewStringUTF)(env, s);
        check();
        return r;
    }
    std::string to_string(jstring s) const {
        if (!s) return std::string();
        const char* p = fn<const char* (*)(JNIEnv*, jstring, jboolean*)>(FN_GetStringUTFChars)(env, s, nullptr);
        std::string out = p ? p : "";
        if (p) fn<void (*)(JNIEnv*, jstring, const char*)>(FN_ReleaseStringUTFChars)(env, s, p);
        return out;
    }
    jobject global(jobject o) const { return o ? fn<jobject (*)(JNIEnv*, jobject)>(FN_NewGlobalRef)(env, o) : nullptr; }
    void drop_global(jobject o) const {
        if (o) fn<void (*)(JNIEnv*, jobject)>(FN_DeleteGlobalRef)(env, o);
    }
    jobject direct_buffer(void* p, jlong n) const {
        jobject r = fn<jobject (*)(JNIEnv*, void*, jlong)>(FN_NewDirectByteBuffer)(env, p, n);
        check();
        return r;
    }
    void* direct_address(jobject buf) const { return fn<void* (*)(JNIEnv*, jobject)>(FN_GetDirectBufferAddress)(env, buf); }
    jint array_length(jobject a) const { return fn<jint (*)(JNIEnv*, jobject)>(FN_GetArrayLength)(env, a); }
    void byte_region(jbyteArray a, jint off, jint n, void* dst) const {
        fn<void (*)(JNIEnv*, jbyteArray, jint, jint, jbyte*)>(FN_GetByteArrayRegion)(env, a, off, n, (jbyte*)dst);
        check();
    }
    jthrowable take_exception() const {   // pending exception -> global ref, cleared on this thread
        jthrowable e = fn<jthrowable (*)(JNIEnv*)>(FN_ExceptionOccurred)(env);
        fn<void (*)(JNIEnv*)>(FN_ExceptionClear)(env);
        return global(e);
    }
};

// local references made inside one upcall sequence die with the frame (native worker threads never return to Java)
struct LocalFrame {
    const J& j;
    explicit LocalFrame(const J& j_) : j(j_) { j.fn<jint (*)(JNIEnv*, jint)>(FN_PushLocalFrame)(j.env, 32); }
    ~LocalFrame() { j.fn<jobject (*)(JNIEnv*, jobject)>(FN_PopLocalFrame)(j.env, nullptr); }
};

std::atomic<uint64_t> g_task_serial{0};

struct JniTask {
    const uint64_t serial = ++g_task_serial;   // identifies the task to threads that outlive it (addresses get reused)
    JavaVM* vm = nullptr;
    auron_task* task = nullptr;
    jobject wrapper = nullptr;         // global ref of AuronCallNativeWrapper (rt.rs:63-73 keeps the same)
    jclass bridge = nullptr;           // global ref of org.apache.auron.jni.JniBridge
    jobject class_loader = nullptr;    // the calling task thread's context class loader and Spark thread context: installed on
    jobject thread_context = nullptr;  // every engine thread that makes upcalls for this task (rt.rs:117-134)
    auron_callbacks cb{};
    std::mutex mu;
    std::map<std::string, jobject> exporters;       // resource id -> AuronArrowFFIExporter
    std::map<std::string, jobject> fs_providers;    // fs resource id -> scala.Function1[String, FileSystem]
    std::map<std::string, jobject> inputs;          // path -> FSDataInputWrapper
    std::map<std::string, jobject> block_iters;     // resource id -> scala.collection.Iterator[BlockObject]
    jobject cur_block = nullptr, cur_buffer = nullptr;
    jobject last_block = nullptr;      // the block handed out last, kept after its close() for throwFetchFailed
    std::string cur_path;
    std::vector<uint8_t> cur_bytes;
    jthrowable failure = nullptr;      // first Java exception raised inside an upcall (any thread)
};

// A native thread that attached itself must detach before it exits (JNI specification, "Detaching from the VM"): the scan's
// producer thread ends with its scan, the pool workers end with the process.  After JniBridge.onExit (a JVM shutdown hook)
// the VM may be gone, so nothing is detached from then on.
std::atomic<bool> g_vm_exiting{false};
struct ThreadAttachment {
    JavaVM* vm = nullptr;
    ~ThreadAttachment() {
        if (vm && !g_vm_exiting.load())
            reinterpret_cast<jint (*)(JavaVM*)>(const_cast<void*>(vm->functions[VM_DetachCurrentThread]))(vm);
    }
};
thread_local ThreadAttachment tl_attachment;

J env_of(JniTask* jt) {
    JNIEnv* env = nullptr;
    auto get_env = reinterpret_cast<jint (*)(JavaVM*, void**, jint)>(const_cast<void*>(jt->vm->functions[VM_GetEnv]));
    if (get_env(jt->vm, (void**)&env, kJniVersion) != 0 || !env) {
        // scan producer / read workers are the engine's own threads (the reference's are tokio workers, rt.rs:117-131)
        auto attach = reinterpret_cast<jint (*)(JavaVM*, void**, void*)>(const_cast<void*>(jt->vm->functions[VM_AttachCurrentThreadAsDaemon]));
        if (attach(jt->vm, (void**)&env, nullptr) == 0) tl_attachment.vm = jt->vm;
    }
    return J{env};
}

// rt.rs:124-134 does this in the worker threads' on_thread_start; pool workers outlive tasks, so it is redone whenever a
// thread first serves another task
thread_local uint64_t tl_context_of = 0;
void adopt_thread_context(const J& j, JniTask* jt) {
    if (tl_context_of == jt->serial || !tl_attachment.vm) return;   // Java threads (the caller of the natives) already carry theirs
    tl_context_of = jt->serial;
    try {
        jvalue a;
        a.l = jt->class_loader;
        j.call_static_void(jt->bridge, "setContextClassLoader", "(Ljava/lang/ClassLoader;)V", &a);
        a.l = jt->thread_context;
        j.call_static_void(jt->bridge, "setThreadContext", "(Ljava/lang/Object;)V", &a);
    } catch (const JavaError&) {
        j.fn<void (*)(JNIEnv*)>(FN_ExceptionClear)(j.env);   // "let _ =" in the reference: best effort
    }
}

// run one upcall sequence; a Java exception is captured for nextBatch to rethrow and reported to the engine as -1
template <typename F>
int64_t upcall(JniTask* jt, F&& body) {
    J j = env_of(jt);
    if (!j.env) return -1;
    LocalFrame frame(j);
    adopt_thread_context(j, jt);
    try {
        return body(j);
    } catch (const JavaError&) {
        jthrowable e = j.take_exception();
        std::lock_guard<std::mutex> g(jt->mu);
        if (!jt->failure) jt->failure = e;
        else j.drop_global(e);
        return -1;
    }
}

jobject get_resource(const J& j, JniTask* jt, const char* id) {
    jvalue a;
    a.l = j.new_string(id);
    return j.call_static_object(jt->bridge, "getResource", "(Ljava/lang/String;)Ljava/lang/Object;", &a);
}

// ---- auron_callbacks over JNI ----------------------------------------------------------------------------------------
int cb_export_next_batch(void* user, const char* resource_id, struct ArrowArray* out) {
    auto* jt = (JniTask*)user;
    return (int)upcall(jt, [&](const J& j) -> int64_t {
        jobject exporter;
        {
            std::lock_guard<std::mutex> g(jt->mu);
            auto it = jt->exporters.find(resource_id);
            if (it == jt->exporters.end()) it = jt->exporters.emplace(resource_id, j.global(get_resource(j, jt, resource_id))).first;
            exporter = it->second;
        }
        if (!exporter) return -1;
        jvalue a;
        a.j = (jlong)(intptr_t)out;
        if (j.call_bool(exporter, "exportNextBatch", "(J)Z", &a)) return 1;
        j.call_void(exporter, "close", "()V");   // ffi_reader_exec.rs:195
        return 0;
    });
}

int64_t cb_read_fully(void* user, const char* fs_resource_id, const char* path, int64_t pos, void* buf, int64_t len) {
    auto* jt = (JniTask*)user;
    return upcall(jt, [&](const J& j) -> int64_t {
        jobject input;
        {
            std::lock_guard<std::mutex> g(jt->mu);
            auto it = jt->inputs.find(path);
            if (it == jt->inputs.end()) {
                auto fp = jt->fs_providers.find(fs_resource_id);
                if (fp == jt->fs_providers.end())
                    fp = jt->fs_providers.emplace(fs_resource_id, j.global(get_resource(j, jt, fs_resource_id))).first;
                if (!fp->second) return -1;   // no such resource: the engine reports the failed read
                jvalue a[2];
                a[0].l = j.new_string(path);
                jobject fs = j.call_object(fp->second, "apply", "(Ljava/lang/Object;)Ljava/lang/Object;", a);   // FsProvider::provide
                a[1].l = a[0].l;
                a[0].l = fs;
                jobject w = j.call_static_object(jt->bridge, "openFileAsDataInputWrapper",
                                                 "(Lorg/apache/hadoop/fs/FileSystem;Ljava/lang/String;)Lorg/apache/auron/hadoop/fs/FSDataInputWrapper;", a);
                it = jt->inputs.emplace(path, j.global(w)).first;
            }
            input = it->second;
        }
        if (!input) return -1;
        jvalue a[2];
        a[0].j = pos;
        a[1].l = j.direct_buffer(buf, len);
        j.call_void(input, "readFully", "(JLjava/nio/ByteBuffer;)V", a);   // throws EOFException on a short read
        return len;
    });
}

// IpcWriterExec's consumer: JniBridge.getResource(id) is a Scala `ByteBuffer => Unit` (ipc_writer_exec.rs:112-118,143-152); every
// delivery wraps the native bytes in a direct ByteBuffer for the duration of the call
int cb_write_ipc(void* user, const char* resource_id, const uint8_t* data, int64_t len) {
    auto* jt = (JniTask*)user;
    return (int)upcall(jt, [&](const J& j) -> int64_t {
        jobject consumer;
        {
            std::lock_guard<std::mutex> g(jt->mu);
            auto it = jt->exporters.find(resource_id);
            if (it == jt->exporters.end()) it = jt->exporters.emplace(resource_id, j.global(get_resource(j, jt, resource_id))).first;
            consumer = it->second;
        }
        if (!consumer) return -1;
        jvalue a;
        a.l = j.direct_buffer(const_cast<uint8_t*>(data), len);
        j.call_object(consumer, "apply", "(Ljava/lang/Object;)Ljava/lang/Object;", &a);
        return 0;
    });
}

// AuronBlockObject.throwFetchFailed(errmsg) on the block handed out last (ipc_reader_exec.rs:211-219; the Java side raises Spark's
// FetchFailedException from it, which stays pending and is rethrown when nextBatch returns)
void cb_fetch_failed(void* user, const char* /*resource_id*/, const char* message) {
    auto* jt = (JniTask*)user;
    upcall(jt, [&](const J& j) -> int64_t {
        jobject block = jt->cur_block ? jt->cur_block : jt->last_block;
        if (!block) return 0;
        jvalue a;
        a.l = j.new_string(message);
        j.call_void(block, "throwFetchFailed", "(Ljava/lang/String;)V", &a);
        return 0;
    });
}

int cb_is_task_running(void* user) {
    auto* jt = (JniTask*)user;
    int64_t r = upcall(jt, [&](const J& j) -> int64_t { return j.call_static_bool(jt->bridge, "isTaskRunning", "()Z") ? 1 : 0; });
    return r > 0;   // an exception while asking counts as "not running" (auron-jni-bridge/src/lib.rs:35-50)
}

// conf.rs:62-116: string entries through JniBridge.stringConf, the integer ones through intConf
int cb_get_conf(void* user, const char* key, char* value, int32_t cap) {
    auto* jt = (JniTask*)user;
    J j = env_of(jt);
    if (!j.env) return -1;
    LocalFrame frame(j);
    try {
        jvalue a;
        a.l = j.new_string(key);
        std::string v;
        if (!strcmp(key, "SPARK_IO_COMPRESSION_CODEC") || !strcmp(key, "SPILL_COMPRESSION_CODEC") || !strcmp(key, "NATIVE_LOG_LEVEL"))
            v = j.to_string((jstring)j.call_static_object(jt->bridge, "stringConf", "(Ljava/lang/String;)Ljava/lang/String;", &a));
        else
            v = std::to_string(j.call_static_int(jt->bridge, "intConf", "(Ljava/lang/String;)I", &a));
        if ((int32_t)v.size() >= cap) return -1;
        memcpy(value, v.c_str(), v.size() + 1);
        return (int)v.size();
    } catch (const JavaError&) {
        j.fn<void (*)(JNIEnv*)>(FN_ExceptionClear)(j.env);   // an entry the JVM side does not know: the engine's default applies
        return -1;
    }
}

void close_current_block(const J& j, JniTask* jt) {
    if (jt->cur_block) {
        jobject b = jt->cur_block;
        jt->cur_block = nullptr;
        j.drop_global(jt->cur_buffer);
        jt->cur_buffer = nullptr;
        j.drop_global(jt->last_block);
        jt->last_block = b;                   // the reader decodes after it has drained the block: a decode failure names this one
        j.call_void(b, "close", "()V");       // the readers close their block when dropped (ipc_reader_exec.rs:383-402)
    }
}

int cb_next_shuffle_block(void* user, const char* resource_id, struct auron_shuffle_block* out) {
    auto* jt = (JniTask*)user;
    return (int)upcall(jt, [&](const J& j) -> int64_t {
        close_current_block(j, jt);
        jobject it;
        {
            std::lock_guard<std::mutex> g(jt->mu);
            auto f = jt->block_iters.find(resource_id);
            if (f == jt->block_iters.end()) {
                jobject provider = get_resource(j, jt, resource_id);
                if (!provider) return -1;
                jobject blocks = j.call_object(provider, "apply", "()Ljava/lang/Object;");   // ipc_reader_exec.rs:150
                f = jt->block_iters.emplace(resource_id, j.global(blocks)).first;
            }
            it = f->second;
        }
        if (!it || !j.call_bool(it, "hasNext", "()Z")) return 0;
        jobject block = j.call_object(it, "next", "()Ljava/lang/Object;");
        jt->cur_block = j.global(block);
        memset(out, 0, sizeof(*out));
        if (j.call_bool(block, "hasFileSegment", "()Z")) {   // get_file_reader, ipc_reader_exec.rs:279-294
            jt->cur_path = j.to_string((jstring)j.call_object(block, "getFilePath", "()Ljava/lang/String;"));
            out->path = jt->cur_path.c_str();
            out->offset = j.call_long(block, "getFileOffset", "()J");
            out->length = j.call_long(block, "getFileLength", "()J");
            return 1;
        }
        if (j.call_bool(block, "hasByteBuffer", "()Z")) {    // get_byte_buffer_reader, :296-307
            jobject bb = j.call_object(block, "getByteBuffer", "()Ljava/nio/ByteBuffer;");
            jint pos = j.call_int(bb, "position", "()I"), remaining = j.call_int(bb, "remaining", "()I");
            if (j.call_bool(bb, "isDirect", "()Z")) {
                jt->cur_buffer = j.global(bb);               // keeps the memory alive until the next call
                out->data = (const uint8_t*)j.direct_address(bb) + pos;
            } else if (j.call_bool(bb, "hasArray", "()Z")) {
                jobject arr = j.call_object(bb, "array", "()Ljava/lang/Object;");
                jint base = j.call_int(bb, "arrayOffset", "()I");
                jt->cur_bytes.resize((size_t)remaining);
                j.byte_region(arr, base + pos, remaining, jt->cur_bytes.data());
                out->data = jt->cur_bytes.data();
            } else {
                return -1;   // "ByteBuffer is not direct and do not have array" (:306)
            }
            out->length = remaining;
            return 1;
        }
        // get_channel_reader (:309-349): drain the ReadableByteChannel into one buffer
        jobject ch = j.call_object(block, "getChannel", "()Ljava/nio/channels/ReadableByteChannel;");
        jt->cur_bytes.clear();
        size_t filled = 0;
        for (;;) {
            const size_t chunk = 1 << 20;
            jt->cur_bytes.resize(filled + chunk);
            LocalFrame inner(j);
            jvalue a;
            a.l = j.direct_buffer(jt->cur_bytes.data() + filled, (jlong)chunk);
            jint got = j.call_int(ch, "read", "(Ljava/nio/ByteBuffer;)I", &a);
            if (got < 0) break;
            filled += (size_t)got;
        }
        j.call_void(ch, "close", "()V");
        jt->cur_bytes.resize(filled);
        out->data = jt->cur_bytes.data();
        out->length = (int64_t)filled;
        return 1;
    });
}

// set_error (rt.rs:309-318): wrapper.setError(new RuntimeException(message, cause)); checkError() rethrows it on the Java side
void set_error(const J& j, JniTask* jt, const char* msg, jthrowable cause) {
    try {
        LocalFrame frame(j);
        jclass rte = j.find_class("java/lang/RuntimeException");
        jvalue a[2];
        a[0].l = j.new_string(msg ? msg : "native execution failed");
        a[1].l = cause;
        jvalue e;
        e.l = j.new_object(rte, "(Ljava/lang/String;Ljava/lang/Throwable;)V", a);
        j.call_void(jt->wrapper, "setError", "(Ljava/lang/Throwable;)V", &e);
    } catch (const JavaError&) {
        // leave whatever the JVM raised pending: the caller of the native sees that instead
    }
}

void throw_runtime(const J& j, const char* msg) {
    if (j.pending()) return;   // a Java exception from an upcall wins
    jclass cls = j.fn<jclass (*)(JNIEnv*, const char*)>(FN_FindClass)(j.env, "java/lang/RuntimeException");
    if (cls) j.fn<jint (*)(JNIEnv*, jclass, const char*)>(FN_ThrowNew)(j.env, cls, msg ? msg : "native execution failed");
}

struct MetricWalk {
    const J* j;
    std::vector<jobject> path;   // MetricNode per depth
};

void release_refs(const J& j, JniTask* jt) {
    for (auto* m : {&jt->exporters, &jt->fs_providers, &jt->block_iters})
        for (auto& kv : *m) j.drop_global(kv.second);
    for (auto& kv : jt->inputs) {
        try {
            if (kv.second) j.call_void(kv.second, "close", "()V");   // FsDataInputWrapper::drop, hadoop_fs.rs:99-105
        } catch (const JavaError&) {
            j.fn<void (*)(JNIEnv*)>(FN_ExceptionClear)(j.env);
        }
        j.drop_global(kv.second);
    }
    j.drop_global(jt->cur_block);
    j.drop_global(jt->last_block);
    j.drop_global(jt->cur_buffer);
    j.drop_global(jt->failure);
    j.drop_global(jt->wrapper);
    j.drop_global(jt->class_loader);
    j.drop_global(jt->thread_context);
    j.drop_global(jt->bridge);
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

// exec.rs:42-118 + rt.rs:75-170.  The memory-overhead and log-level arguments configure the Rust memory manager and
// logger; the GPU engine sizes itself from the device and logs through stderr, so they are accepted and unused.
jlong Java_org_apache_auron_jni_JniBridge_callNative(JNIEnv* env, jclass, jlong /*executor_memory_overhead*/, jstring /*log_level*/,
                                                     jobject native_wrapper) {
    J j{env};
    auto* jt = new JniTask;
    try {
        j.fn<jint (*)(JNIEnv*, JavaVM**)>(FN_GetJavaVM)(env, &jt->vm);
        jt->wrapper = j.global(native_wrapper);
        jt->bridge = (jclass)j.global(j.find_class("org/apache/auron/jni/JniBridge"));
        jt->thread_context = j.global(j.call_static_object(jt->bridge, "getThreadContext", "()Ljava/lang/Object;"));          // rt.rs:117-121
        jt->class_loader = j.global(j.call_static_object(jt->bridge, "getContextClassLoader", "()Ljava/lang/ClassLoader;"));
        jbyteArray arr = (jbyteArray)j.call_object(native_wrapper, "getRawTaskDefinition", "()[B");   // rt.rs:78-83
        jint n = j.array_length(arr);
        std::vector<uint8_t> bytes((size_t)n);
        j.byte_region(arr, 0, n, bytes.data());
        jt->cb.user = jt;
        jt->cb.export_next_batch = cb_export_next_batch;
        jt->cb.read_fully = getenv("AURON_B200_LOCAL_FS") ? nullptr : cb_read_fully;   // escape hatch: read paths from the local FS
        jt->cb.is_task_running = cb_is_task_running;
        jt->cb.next_shuffle_block = cb_next_shuffle_block;
        jt->cb.upcalls_from_any_thread = 1;
        jt->cb.get_conf = cb_get_conf;
        jt->cb.write_ipc = cb_write_ipc;
        jt->cb.fetch_failed = cb_fetch_failed;
        const char* dev = getenv("AURON_B200_DEVICE");
        jt->task = auron_b200_call_native(bytes.data(), bytes.size(), &jt->cb, dev ? atoi(dev) : 0);
        if (!jt->task) {
            throw_runtime(j, auron_b200_last_error());
            release_refs(j, jt);
            delete jt;
            return 0;
        }
        // importSchema(ffiSchemaPtr) (rt.rs:167-170): the Java side wraps and releases the struct
        ArrowSchema schema;
        memset(&schema, 0, sizeof(schema));
        if (auron_b200_schema(jt->task, &schema) != 0) {
            throw_runtime(j, auron_b200_last_error());
        } else {
            jvalue a;
            a.j = (jlong)(intptr_t)&schema;
            try {
                j.call_void(native_wrapper, "importSchema", "(J)V", &a);
            } catch (const JavaError&) {
            }
            if (schema.release) schema.release(&schema);
        }
        if (j.pending()) {
            auron_b200_finalize_native(jt->task);
            jthrowable pending = j.fn<jthrowable (*)(JNIEnv*)>(FN_ExceptionOccurred)(env);
            j.fn<void (*)(JNIEnv*)>(FN_ExceptionClear)(env);
            release_refs(j, jt);
            j.fn<jint (*)(JNIEnv*, jthrowable)>(FN_Throw)(env, pending);
            delete jt;
            return 0;
        }
        return (jlong)(intptr_t)jt;
    } catch (const JavaError&) {
        // the Java exception stays pending for the caller
        if (jt->task) auron_b200_finalize_native(jt->task);
        release_refs(j, jt);
        delete jt;
        return 0;
    }
}

// exec.rs:122-129 + rt.rs:250-280: deliver the next batch through wrapper.importBatch(long ffiArrayPtr)
jboolean Java_org_apache_auron_jni_JniBridge_nextBatch(JNIEnv* env, jclass, jlong ptr) {
    auto* jt = reinterpret_cast<JniTask*>((intptr_t)ptr);
    if (!jt || !jt->task) return 0;
    J j{env};
    ArrowArray arr;
    memset(&arr, 0, sizeof(arr));
    int rc = auron_b200_next_batch(jt->task, &arr);
    if (rc < 0) {
        jthrowable cause;
        {
            std::lock_guard<std::mutex> g(jt->mu);
            cause = jt->failure;
            jt->failure = nullptr;
        }
        set_error(j, jt, auron_b200_last_error(), cause);
        j.drop_global(cause);
        return 0;
    }
    if (rc == 0) return 0;
    jvalue a;
    a.j = (jlong)(intptr_t)&arr;
    try {
        j.call_void(jt->wrapper, "importBatch", "(J)V", &a);   // rt.rs:258-262; the Java side moves the array out
    } catch (const JavaError&) {
        if (arr.release) arr.release(&arr);
        return 0;   // exception pending for the caller
    }
    if (arr.release) arr.release(&arr);
    return 1;
}

// exec.rs:133-140 + rt.rs:282-306: push the metrics into the wrapper's MetricNode tree, then tear down
void Java_org_apache_auron_jni_JniBridge_finalizeNative(JNIEnv* env, jclass, jlong ptr) {
    auto* jt = reinterpret_cast<JniTask*>((intptr_t)ptr);
    if (!jt) return;
    J j{env};
    try {
        LocalFrame frame(j);
        jobject root = j.call_object(jt->wrapper, "getMetrics", "()Lorg/apache/auron/metric/MetricNode;");
        if (root && jt->task) {
            MetricWalk w{&j, {}};
            auto enter = [](void* user, int depth, int child_index, const char*) {
                auto* mw = (MetricWalk*)user;
                mw->path.resize((size_t)depth + 1);
                if (depth == 0) return;   // root installed by the caller
                jobject parent = mw->path[(size_t)depth - 1];
                jobject node = nullptr;
                if (parent) {
                    try {
                        jvalue a;
                        a.i = child_index;
                        node = mw->j->call_object(parent, "getChild", "(I)Lorg/apache/auron/metric/MetricNode;", &a);
                    } catch (const JavaError&) {
                        mw->j->fn<void (*)(JNIEnv*)>(FN_ExceptionClear)(mw->j->env);
                    }
                }
                mw->path[(size_t)depth] = node;
            };
            auto metric = [](void* user, int depth, const char*, const char* name, int64_t value) {
                auto* mw = (MetricWalk*)user;
                if (depth < 0 || (size_t)depth >= mw->path.size() || !mw->path[(size_t)depth]) return;
                try {
                    jvalue a[2];
                    a[0].l = mw->j->new_string(name);
                    a[1].j = value;
                    mw->j->call_void(mw->path[(size_t)depth], "add", "(Ljava/lang/String;J)V", a);
                } catch (const JavaError&) {
                    mw->j->fn<void (*)(JNIEnv*)>(FN_ExceptionClear)(mw->j->env);
                }
            };
            w.path.push_back(root);
            auron_b200_metrics_walk(jt->task, enter, metric, &w);
        }
    } catch (const JavaError&) {
        j.fn<void (*)(JNIEnv*)>(FN_ExceptionClear)(env);   // update_metrics().unwrap_or_default() (rt.rs:286)
    }
    auron_b200_finalize_native(jt->task);   // joins the scan producer: no upcall can be in flight after this
    jt->task = nullptr;
    try {
        close_current_block(j, jt);
    } catch (const JavaError&) {
        j.fn<void (*)(JNIEnv*)>(FN_ExceptionClear)(env);
    }
    release_refs(j, jt);
    delete jt;
}

// exec.rs:144-149
void Java_org_apache_auron_jni_JniBridge_onExit(JNIEnv*, jclass) {
    g_vm_exiting.store(true);
    auron_b200_on_exit();
}

}  // extern "C"
#pragma GCC visibility pop
