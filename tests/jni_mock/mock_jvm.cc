// mock_jvm.cc -- TEST INFRASTRUCTURE.  A stand-in for the JVM side of Apache Auron's JNI boundary, so that the JNI face of
// libauron_b200.so (auron_b200/csrc/jni_face.cc) can be driven end to end without a JVM (none exists in this image).
//
// It provides a JNIEnv / JavaVM function table (laid out by index from the JNI specification; every slot the product does not
// need traps with its index) and the Java classes the reference's native side talks to on this path, with the same
// names, method names and signatures (native-engine/auron-jni-bridge/src/jni_bridge.rs):
//
//   org.apache.auron.jni.AuronCallNativeWrapper   getRawTaskDefinition()[B importSchema(J)V importBatch(J)V
//                                                 setError(Ljava/lang/Throwable;)V getMetrics()
//   org.apache.auron.jni.JniBridge (static)       getResource isTaskRunning openFileAsDataInputWrapper
//   org.apache.auron.arrowio.AuronArrowFFIExporter exportNextBatch(J)Z close()V
//   scala.Function0 / Function1 / collection.Iterator, org.apache.spark.sql.execution.auron.shuffle.BlockObject,
//   java.nio.ByteBuffer (direct and heap), java.nio.channels.ReadableByteChannel,
//   org.apache.auron.hadoop.fs.FSDataInputWrapper readFully(JLjava/nio/ByteBuffer;)V close()V,
//   org.apache.auron.metric.MetricNode getChild(I) add(Ljava/lang/String;J)V, java.lang.RuntimeException(String, Throwable)
//
// A method looked up with a name or signature that the real class does not have raises NoSuchMethodError, as a JVM would.
// Reference counting of global references and local frames is tracked so the tests can assert that nothing leaks.
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <utility>
#include <vector>

namespace {

// Arrow C data interface structs (https://arrow.apache.org/docs/format/CDataInterface.html)
struct ArrowSchema {
    const char* format;
    const char* name;
    const char* metadata;
    int64_t flags;
    int64_t n_children;
    ArrowSchema** children;
    ArrowSchema* dictionary;
    void (*release)(ArrowSchema*);
    void* private_data;
};
struct ArrowArray {
    int64_t length, null_count, offset, n_buffers, n_children;
    const void** buffers;
    ArrowArray** children;
    ArrowArray* dictionary;
    void (*release)(ArrowArray*);
    void* private_data;
};

union jvalue {
    uint8_t z;
    int8_t b;
    uint16_t c;
    int16_t s;
    int32_t i;
    int64_t j;
    float f;
    double d;
    void* l;
};

struct Jvm;
struct Obj {
    std::string cls;   // binary name with '/' separators
    virtual ~Obj() {}
};
struct ClassObj : Obj {
    std::string name;
};
struct StringObj : Obj {
    std::string s;
};
struct ByteArrayObj : Obj {
    std::vector<uint8_t> bytes;
};
struct ThrowableObj : Obj {
    std::string message;
    ThrowableObj* cause = nullptr;
};
struct MetricNodeObj : Obj {
    std::map<int, MetricNodeObj*> children;
    std::vector<std::pair<std::string, int64_t>> values;
};
struct WrapperObj : Obj {
    ByteArrayObj* task_definition = nullptr;
    MetricNodeObj* metrics = nullptr;
    ThrowableObj* error = nullptr;
    bool have_schema = false;
    ArrowSchema schema{};
    std::vector<ArrowArray> batches;
    int fail_import_after = -1;   // importBatch throws once this many batches were imported
};
typedef int (*export_fn)(void* user, void* out_array);   // 1 = batch written, 0 = end, <0 = throw
struct ExporterObj : Obj {
    export_fn fn = nullptr;
    void* user = nullptr;
    int closed = 0;
};
struct FsProviderObj : Obj {};
struct FileSystemObj : Obj {};
struct InputWrapperObj : Obj {
    int fd = -1;
    int closed = 0;
    std::atomic<int64_t> reads{0};
};
struct ByteBufferObj : Obj {
    bool direct = false;
    uint8_t* address = nullptr;   // direct
    int64_t capacity = 0;
    ByteArrayObj* array = nullptr;   // heap
    int array_offset = 0, position = 0, limit = 0;
};
struct ChannelObj : Obj {
    std::vector<uint8_t> bytes;
    size_t pos = 0;
    int closed = 0;
    size_t max_read = 1 << 16;   // a channel may return short reads
};
struct BlockObj : Obj {
    int kind = 0;   // 0 file segment, 1 direct buffer, 2 heap buffer, 3 channel
    std::string path;
    int64_t offset = 0, length = 0;
    ByteBufferObj* buffer = nullptr;
    ChannelObj* channel = nullptr;
    int closed = 0;
};
struct IteratorObj : Obj {
    std::vector<BlockObj*> blocks;
    size_t next = 0;
};
struct BlocksProviderObj : Obj {
    IteratorObj* it = nullptr;
};
struct Method {
    std::string cls, name, sig;
    bool is_static;
};

struct Env {   // JNIEnv: first member is the function table pointer
    void* const* functions;
    Jvm* vm;
    Obj* class_loader = nullptr;     // per-thread state of the Java side
    Obj* thread_context = nullptr;
    ThrowableObj* pending = nullptr;
    int frames = 0;
};
struct Vm {   // JavaVM
    void* const* functions;
    Jvm* jvm;
};

struct Jvm {
    Vm vm;
    std::mutex mu;
    std::vector<std::unique_ptr<Obj>> heap;   // everything lives until mock_free
    std::map<std::string, ClassObj*> classes;
    std::vector<std::unique_ptr<Method>> methods;
    std::vector<std::unique_ptr<Env>> envs;
    std::map<std::string, Obj*> resources;
    std::map<std::string, std::string> conf;
    std::multiset<void*> globals;
    std::atomic<int64_t> reads_without_context{0};   // readFully on a thread that carries no class loader / thread context
    std::atomic<int64_t> global_new{0}, global_del{0}, attached{0}, detached{0}, frames_pushed{0}, frames_popped{0};
    bool task_running = true;
    std::string trouble;   // protocol violations noticed by the mock

    template <typename T>
    T* make(const std::string& cls) {
        auto* o = new T;
        o->cls = cls;
        std::lock_guard<std::mutex> g(mu);
        heap.emplace_back(o);
        return o;
    }
    ClassObj* klass(const std::string& name) {
        std::lock_guard<std::mutex> g(mu);
        auto it = classes.find(name);
        if (it != classes.end()) return it->second;
        auto* c = new ClassObj;
        c->cls = "java/lang/Class";
        c->name = name;
        heap.emplace_back(c);
        classes[name] = c;
        return c;
    }
    void complain(const std::string& s) {
        std::lock_guard<std::mutex> g(mu);
        trouble += s + "\n";
    }
};

thread_local Env* tl_env = nullptr;
thread_local Jvm* tl_env_owner = nullptr;

void raise(Env* e, const char* cls, const std::string& msg) {
    auto* t = e->vm->make<ThrowableObj>(cls);
    t->message = msg;
    e->pending = t;
}

// the methods each mock class really has: (class, name, signature)
const char* kMethods[][3] = {
    {"org/apache/auron/jni/AuronCallNativeWrapper", "getRawTaskDefinition", "()[B"},
    {"org/apache/auron/jni/AuronCallNativeWrapper", "importSchema", "(J)V"},
    {"org/apache/auron/jni/AuronCallNativeWrapper", "importBatch", "(J)V"},
    {"org/apache/auron/jni/AuronCallNativeWrapper", "setError", "(Ljava/lang/Throwable;)V"},
    {"org/apache/auron/jni/AuronCallNativeWrapper", "getMetrics", "()Lorg/apache/auron/metric/MetricNode;"},
    {"org/apache/auron/jni/JniBridge", "getResource", "(Ljava/lang/String;)Ljava/lang/Object;"},
    {"org/apache/auron/jni/JniBridge", "isTaskRunning", "()Z"},
    {"org/apache/auron/jni/JniBridge", "openFileAsDataInputWrapper",
     "(Lorg/apache/hadoop/fs/FileSystem;Ljava/lang/String;)Lorg/apache/auron/hadoop/fs/FSDataInputWrapper;"},
    {"org/apache/auron/jni/JniBridge", "getContextClassLoader", "()Ljava/lang/ClassLoader;"},
    {"org/apache/auron/jni/JniBridge", "setContextClassLoader", "(Ljava/lang/ClassLoader;)V"},
    {"org/apache/auron/jni/JniBridge", "getThreadContext", "()Ljava/lang/Object;"},
    {"org/apache/auron/jni/JniBridge", "setThreadContext", "(Ljava/lang/Object;)V"},
    {"org/apache/auron/jni/JniBridge", "stringConf", "(Ljava/lang/String;)Ljava/lang/String;"},
    {"org/apache/auron/jni/JniBridge", "intConf", "(Ljava/lang/String;)I"},
    {"org/apache/auron/arrowio/AuronArrowFFIExporter", "exportNextBatch", "(J)Z"},
    {"org/apache/auron/arrowio/AuronArrowFFIExporter", "close", "()V"},
    {"scala/Function0", "apply", "()Ljava/lang/Object;"},
    {"scala/Function1", "apply", "(Ljava/lang/Object;)Ljava/lang/Object;"},
    {"scala/collection/Iterator", "hasNext", "()Z"},
    {"scala/collection/Iterator", "next", "()Ljava/lang/Object;"},
    {"org/apache/spark/sql/execution/auron/shuffle/BlockObject", "hasFileSegment", "()Z"},
    {"org/apache/spark/sql/execution/auron/shuffle/BlockObject", "hasByteBuffer", "()Z"},
    {"org/apache/spark/sql/execution/auron/shuffle/BlockObject", "getFilePath", "()Ljava/lang/String;"},
    {"org/apache/spark/sql/execution/auron/shuffle/BlockObject", "getFileOffset", "()J"},
    {"org/apache/spark/sql/execution/auron/shuffle/BlockObject", "getFileLength", "()J"},
    {"org/apache/spark/sql/execution/auron/shuffle/BlockObject", "getByteBuffer", "()Ljava/nio/ByteBuffer;"},
    {"org/apache/spark/sql/execution/auron/shuffle/BlockObject", "getChannel", "()Ljava/nio/channels/ReadableByteChannel;"},
    {"org/apache/spark/sql/execution/auron/shuffle/BlockObject", "throwFetchFailed", "(Ljava/lang/String;)V"},
    {"org/apache/spark/sql/execution/auron/shuffle/BlockObject", "close", "()V"},
    {"java/nio/ByteBuffer", "isDirect", "()Z"},
    {"java/nio/ByteBuffer", "hasArray", "()Z"},
    {"java/nio/ByteBuffer", "array", "()Ljava/lang/Object;"},
    {"java/nio/ByteBuffer", "arrayOffset", "()I"},
    {"java/nio/ByteBuffer", "position", "()I"},
    {"java/nio/ByteBuffer", "remaining", "()I"},
    {"java/nio/ByteBuffer", "hasRemaining", "()Z"},
    {"java/nio/channels/ReadableByteChannel", "read", "(Ljava/nio/ByteBuffer;)I"},
    {"java/nio/channels/ReadableByteChannel", "close", "()V"},
    {"org/apache/auron/hadoop/fs/FSDataInputWrapper", "readFully", "(JLjava/nio/ByteBuffer;)V"},
    {"org/apache/auron/hadoop/fs/FSDataInputWrapper", "close", "()V"},
    {"org/apache/auron/metric/MetricNode", "getChild", "(I)Lorg/apache/auron/metric/MetricNode;"},
    {"org/apache/auron/metric/MetricNode", "add", "(Ljava/lang/String;J)V"},
    {"java/lang/RuntimeException", "<init>", "(Ljava/lang/String;Ljava/lang/Throwable;)V"},
    {"java/lang/RuntimeException", "<init>", "(Ljava/lang/String;)V"},
};

Method* lookup(Env* e, ClassObj* c, const char* name, const char* sig, bool is_static) {
    if (!c) {
        raise(e, "java/lang/NullPointerException", "class is null");
        return nullptr;
    }
    bool known = false;
    for (auto& m : kMethods)
        if (c->name == m[0] && !strcmp(name, m[1]) && !strcmp(sig, m[2])) known = true;
    if (!known) {
        raise(e, "java/lang/NoSuchMethodError", c->name + "." + name + sig);
        return nullptr;
    }
    auto* m = new Method{c->name, name, sig, is_static};
    std::lock_guard<std::mutex> g(e->vm->mu);
    e->vm->methods.emplace_back(m);
    return m;
}

// ---- the behaviour of the mock classes --------------------------------------------------------------------------------
jvalue invoke(Env* e, Obj* self, Method* m, const jvalue* a) {
    jvalue r;
    r.j = 0;
    Jvm* vm = e->vm;
    if (e->pending) vm->complain("JNI call of " + m->name + " with an exception pending");
    if (!m->is_static && !self) {
        raise(e, "java/lang/NullPointerException", m->name);
        return r;
    }
    const std::string& n = m->name;
    if (m->cls == "org/apache/auron/jni/JniBridge") {
        if (n == "getResource") {
            auto* key = (StringObj*)a[0].l;
            std::lock_guard<std::mutex> g(vm->mu);
            auto it = vm->resources.find(key->s);
            r.l = it == vm->resources.end() ? nullptr : it->second;
        } else if (n == "getContextClassLoader") {
            r.l = e->class_loader;
        } else if (n == "getThreadContext") {
            r.l = e->thread_context;
        } else if (n == "setContextClassLoader") {
            e->class_loader = (Obj*)a[0].l;
        } else if (n == "setThreadContext") {
            e->thread_context = (Obj*)a[0].l;
        } else if (n == "stringConf" || n == "intConf") {
            auto* key = (StringObj*)a[0].l;
            std::string v;
            bool found;
            {
                std::lock_guard<std::mutex> g(vm->mu);
                auto it = vm->conf.find(key->s);
                found = it != vm->conf.end();
                if (found) v = it->second;
            }
            if (!found) {
                raise(e, "java/util/NoSuchElementException", key->s);
                return r;
            }
            if (n == "intConf") r.i = atoi(v.c_str());
            else {
                auto* s = vm->make<StringObj>("java/lang/String");
                s->s = v;
                r.l = s;
            }
        } else if (n == "isTaskRunning") {
            r.z = vm->task_running;
        } else if (n == "openFileAsDataInputWrapper") {
            auto* fs = (Obj*)a[0].l;
            auto* path = (StringObj*)a[1].l;
            if (!fs || fs->cls != "org/apache/hadoop/fs/FileSystem") {
                raise(e, "java/lang/ClassCastException", "not a FileSystem");
                return r;
            }
            std::string p = path->s.rfind("file://", 0) == 0 ? path->s.substr(7) : path->s;
            int fd = open(p.c_str(), O_RDONLY);
            if (fd < 0) {
                raise(e, "java/io/FileNotFoundException", p);
                return r;
            }
            auto* w = vm->make<InputWrapperObj>("org/apache/auron/hadoop/fs/FSDataInputWrapper");
            w->fd = fd;
            std::lock_guard<std::mutex> g(vm->mu);
            vm->resources["__input__:" + p + ":" + std::to_string((uintptr_t)w)] = w;
            r.l = w;
        }
        return r;
    }
    if (auto* w = dynamic_cast<WrapperObj*>(self)) {
        if (n == "getRawTaskDefinition") r.l = w->task_definition;
        else if (n == "getMetrics") r.l = w->metrics;
        else if (n == "setError") w->error = (ThrowableObj*)a[0].l;
        else if (n == "importSchema") {   // ArrowSchema.wrap(ptr) + Data.importSchema: takes the struct over, source released
            auto* s = (ArrowSchema*)(intptr_t)a[0].j;
            if (w->have_schema && w->schema.release) w->schema.release(&w->schema);
            w->schema = *s;
            s->release = nullptr;
            w->have_schema = true;
        } else if (n == "importBatch") {
            if (w->fail_import_after >= 0 && (int)w->batches.size() >= w->fail_import_after) {
                raise(e, "java/lang/IllegalStateException", "consumer failed");
                return r;
            }
            auto* arr = (ArrowArray*)(intptr_t)a[0].j;
            w->batches.push_back(*arr);
            arr->release = nullptr;
        }
        return r;
    }
    if (auto* x = dynamic_cast<ExporterObj*>(self)) {
        if (n == "close") x->closed++;
        else {
            int rc = x->fn(x->user, (void*)(intptr_t)a[0].j);
            if (rc < 0) raise(e, "java/lang/IllegalStateException", "exporter failed");
            r.z = rc > 0;
        }
        return r;
    }
    if (dynamic_cast<FsProviderObj*>(self)) {
        r.l = vm->make<FileSystemObj>("org/apache/hadoop/fs/FileSystem");
        return r;
    }
    if (auto* p = dynamic_cast<BlocksProviderObj*>(self)) {
        r.l = p->it;
        return r;
    }
    if (auto* it = dynamic_cast<IteratorObj*>(self)) {
        if (n == "hasNext") r.z = it->next < it->blocks.size();
        else if (it->next < it->blocks.size()) r.l = it->blocks[it->next++];
        else raise(e, "java/util/NoSuchElementException", "next on empty iterator");
        return r;
    }
    if (auto* b = dynamic_cast<BlockObj*>(self)) {
        if (n == "hasFileSegment") r.z = b->kind == 0;
        else if (n == "hasByteBuffer") r.z = b->kind == 1 || b->kind == 2;
        else if (n == "getFilePath") {
            auto* s = vm->make<StringObj>("java/lang/String");
            s->s = b->path;
            r.l = s;
        } else if (n == "getFileOffset") r.j = b->offset;
        else if (n == "getFileLength") r.j = b->length;
        else if (n == "getByteBuffer") r.l = b->buffer;
        else if (n == "getChannel") r.l = b->channel;
        else if (n == "close") b->closed++;
        else if (n == "throwFetchFailed") raise(e, "org/apache/spark/shuffle/FetchFailedException", ((StringObj*)a[0].l)->s);
        return r;
    }
    if (auto* bb = dynamic_cast<ByteBufferObj*>(self)) {
        if (n == "isDirect") r.z = bb->direct;
        else if (n == "hasArray") r.z = !bb->direct;
        else if (n == "array") r.l = bb->array;
        else if (n == "arrayOffset") r.i = bb->array_offset;
        else if (n == "position") r.i = bb->position;
        else if (n == "remaining") r.i = bb->limit - bb->position;
        else if (n == "hasRemaining") r.z = bb->limit > bb->position;
        return r;
    }
    if (auto* ch = dynamic_cast<ChannelObj*>(self)) {
        if (n == "close") ch->closed++;
        else {
            auto* dst = (ByteBufferObj*)a[0].l;
            if (ch->pos >= ch->bytes.size()) {
                r.i = -1;
                return r;
            }
            size_t room = (size_t)(dst->limit - dst->position);
            size_t take = std::min(std::min(room, ch->bytes.size() - ch->pos), ch->max_read);
            memcpy(dst->address + dst->position, ch->bytes.data() + ch->pos, take);
            dst->position += (int)take;
            ch->pos += take;
            r.i = (int32_t)take;
        }
        return r;
    }
    if (auto* in = dynamic_cast<InputWrapperObj*>(self)) {
        if (n == "close") in->closed++;
        else {
            auto* dst = (ByteBufferObj*)a[1].l;
            int64_t pos = a[0].j, want = dst->limit - dst->position, done = 0;
            while (done < want) {
                ssize_t got = pread(in->fd, dst->address + dst->position + done, (size_t)(want - done), pos + done);
                if (got <= 0) break;
                done += got;
            }
            in->reads++;
            if (!e->class_loader || !e->thread_context) vm->reads_without_context++;
            if (done < want) raise(e, "java/io/EOFException", "cannot read more " + std::to_string(want - done) + " bytes");
        }
        return r;
    }
    if (auto* mn = dynamic_cast<MetricNodeObj*>(self)) {
        if (n == "getChild") {
            std::lock_guard<std::mutex> g(vm->mu);
            auto it = mn->children.find(a[0].i);
            r.l = it == mn->children.end() ? nullptr : it->second;   // MetricNode.getChild returns null past the known children
        } else {
            mn->values.emplace_back(((StringObj*)a[0].l)->s, a[1].j);
        }
        return r;
    }
    vm->complain("call of " + m->cls + "." + m->name + " on an object of class " + self->cls);
    raise(e, "java/lang/IncompatibleClassChangeError", m->name);
    return r;
}

// ---- JNIEnv functions -------------------------------------------------------------------------------------------------
template <int N>
void trap() {
    fprintf(stderr, "mock_jvm: JNI function table slot %d is not provided by the mock\n", N);
    abort();
}
template <int... I>
void fill_traps(void** t, std::integer_sequence<int, I...>) {
    ((t[I] = (void*)&trap<I>), ...);
}

void* f_FindClass(Env* e, const char* name) { return e->vm->klass(name); }
int32_t f_Throw(Env* e, void* t) {
    e->pending = (ThrowableObj*)t;
    return 0;
}
int32_t f_ThrowNew(Env* e, void* cls, const char* msg) {
    raise(e, ((ClassObj*)cls)->name.c_str(), msg ? msg : "");
    return 0;
}
void* f_ExceptionOccurred(Env* e) { return e->pending; }
void f_ExceptionClear(Env* e) { e->pending = nullptr; }
int32_t f_PushLocalFrame(Env* e, int32_t) {
    e->frames++;
    e->vm->frames_pushed++;
    return 0;
}
void* f_PopLocalFrame(Env* e, void* result) {
    if (--e->frames < 0) e->vm->complain("PopLocalFrame without PushLocalFrame");
    e->vm->frames_popped++;
    return result;
}
void* f_NewGlobalRef(Env* e, void* o) {
    if (!o) return nullptr;
    std::lock_guard<std::mutex> g(e->vm->mu);
    e->vm->globals.insert(o);
    e->vm->global_new++;
    return o;
}
void f_DeleteGlobalRef(Env* e, void* o) {
    if (!o) return;
    std::lock_guard<std::mutex> g(e->vm->mu);
    auto it = e->vm->globals.find(o);
    if (it == e->vm->globals.end()) {
        e->vm->trouble += "DeleteGlobalRef of a reference that is not a live global\n";
        return;
    }
    e->vm->globals.erase(it);
    e->vm->global_del++;
}
void* f_NewObjectA(Env* e, void* cls, Method* m, const jvalue* a) {
    if (!m) return nullptr;
    auto* t = e->vm->make<ThrowableObj>(((ClassObj*)cls)->name);
    t->message = a[0].l ? ((StringObj*)a[0].l)->s : "";
    if (m->sig == "(Ljava/lang/String;Ljava/lang/Throwable;)V") t->cause = (ThrowableObj*)a[1].l;
    return t;
}
void* f_GetObjectClass(Env* e, Obj* o) { return o ? e->vm->klass(o->cls) : nullptr; }
void* f_GetMethodID(Env* e, void* cls, const char* name, const char* sig) { return lookup(e, (ClassObj*)cls, name, sig, false); }
void* f_GetStaticMethodID(Env* e, void* cls, const char* name, const char* sig) { return lookup(e, (ClassObj*)cls, name, sig, true); }
void* f_CallObjectMethodA(Env* e, Obj* o, Method* m, const jvalue* a) { return m ? invoke(e, o, m, a).l : nullptr; }
uint8_t f_CallBooleanMethodA(Env* e, Obj* o, Method* m, const jvalue* a) { return m ? invoke(e, o, m, a).z : 0; }
int32_t f_CallIntMethodA(Env* e, Obj* o, Method* m, const jvalue* a) { return m ? invoke(e, o, m, a).i : 0; }
int64_t f_CallLongMethodA(Env* e, Obj* o, Method* m, const jvalue* a) { return m ? invoke(e, o, m, a).j : 0; }
void f_CallVoidMethodA(Env* e, Obj* o, Method* m, const jvalue* a) {
    if (m) invoke(e, o, m, a);
}
void* f_CallStaticObjectMethodA(Env* e, void*, Method* m, const jvalue* a) { return m ? invoke(e, nullptr, m, a).l : nullptr; }
uint8_t f_CallStaticBooleanMethodA(Env* e, void*, Method* m, const jvalue* a) { return m ? invoke(e, nullptr, m, a).z : 0; }
void* f_NewStringUTF(Env* e, const char* s) {
    auto* o = e->vm->make<StringObj>("java/lang/String");
    o->s = s ? s : "";
    return o;
}
const char* f_GetStringUTFChars(Env*, StringObj* s, uint8_t* is_copy) {
    if (is_copy) *is_copy = 1;
    return strdup(s->s.c_str());
}
void f_ReleaseStringUTFChars(Env*, StringObj*, const char* p) { free((void*)p); }
int32_t f_GetArrayLength(Env*, ByteArrayObj* a) { return (int32_t)a->bytes.size(); }
void f_GetByteArrayRegion(Env* e, ByteArrayObj* a, int32_t off, int32_t n, int8_t* dst) {
    if (off < 0 || n < 0 || (size_t)off + (size_t)n > a->bytes.size()) {
        raise(e, "java/lang/ArrayIndexOutOfBoundsException", "GetByteArrayRegion");
        return;
    }
    memcpy(dst, a->bytes.data() + off, (size_t)n);
}
void f_CallStaticVoidMethodA(Env* e, void*, Method* m, const jvalue* a) {
    if (m) invoke(e, nullptr, m, a);
}
int32_t f_CallStaticIntMethodA(Env* e, void*, Method* m, const jvalue* a) { return m ? invoke(e, nullptr, m, a).i : 0; }
int32_t f_GetJavaVM(Env* e, void** out) {
    *out = &e->vm->vm;
    return 0;
}
uint8_t f_ExceptionCheck(Env* e) { return e->pending != nullptr; }
void* f_NewDirectByteBuffer(Env* e, void* addr, int64_t cap) {
    auto* bb = e->vm->make<ByteBufferObj>("java/nio/ByteBuffer");
    bb->direct = true;
    bb->address = (uint8_t*)addr;
    bb->capacity = cap;
    bb->limit = (int)cap;
    return bb;
}
void* f_GetDirectBufferAddress(Env*, ByteBufferObj* bb) { return bb && bb->direct ? bb->address : nullptr; }

void* g_table[240];
void* g_vm_table[8];
std::once_flag g_once;

Env* new_env(Jvm* vm) {
    auto* e = new Env;
    e->functions = g_table;
    e->vm = vm;
    std::lock_guard<std::mutex> g(vm->mu);
    vm->envs.emplace_back(e);
    return e;
}
int32_t vm_GetEnv(Vm* v, void** out, int32_t) {
    if (tl_env && tl_env_owner == v->jvm) {
        *out = tl_env;
        return 0;
    }
    *out = nullptr;
    return -2;   // JNI_EDETACHED
}
int32_t vm_Attach(Vm* v, void** out, void*) {
    if (!(tl_env && tl_env_owner == v->jvm)) {
        tl_env = new_env(v->jvm);
        tl_env_owner = v->jvm;
        v->jvm->attached++;
    }
    *out = tl_env;
    return 0;
}

int32_t vm_Detach(Vm* v) {
    if (tl_env && tl_env_owner == v->jvm) {
        if (tl_env->frames != 0) v->jvm->complain("DetachCurrentThread with an open local frame");
        tl_env = nullptr;
        tl_env_owner = nullptr;
        v->jvm->detached++;
        return 0;
    }
    return -2;
}

void init_tables() {
    fill_traps(g_table, std::make_integer_sequence<int, 240>());
    for (int i = 0; i < 4; i++) g_table[i] = nullptr;   // reserved slots
    g_table[6] = (void*)f_FindClass;
    g_table[13] = (void*)f_Throw;
    g_table[14] = (void*)f_ThrowNew;
    g_table[15] = (void*)f_ExceptionOccurred;
    g_table[17] = (void*)f_ExceptionClear;
    g_table[19] = (void*)f_PushLocalFrame;
    g_table[20] = (void*)f_PopLocalFrame;
    g_table[21] = (void*)f_NewGlobalRef;
    g_table[22] = (void*)f_DeleteGlobalRef;
    g_table[30] = (void*)f_NewObjectA;
    g_table[31] = (void*)f_GetObjectClass;
    g_table[33] = (void*)f_GetMethodID;
    g_table[36] = (void*)f_CallObjectMethodA;
    g_table[39] = (void*)f_CallBooleanMethodA;
    g_table[51] = (void*)f_CallIntMethodA;
    g_table[54] = (void*)f_CallLongMethodA;
    g_table[63] = (void*)f_CallVoidMethodA;
    g_table[113] = (void*)f_GetStaticMethodID;
    g_table[116] = (void*)f_CallStaticObjectMethodA;
    g_table[119] = (void*)f_CallStaticBooleanMethodA;
    g_table[131] = (void*)f_CallStaticIntMethodA;
    g_table[143] = (void*)f_CallStaticVoidMethodA;
    g_table[167] = (void*)f_NewStringUTF;
    g_table[169] = (void*)f_GetStringUTFChars;
    g_table[170] = (void*)f_ReleaseStringUTFChars;
    g_table[171] = (void*)f_GetArrayLength;
    g_table[200] = (void*)f_GetByteArrayRegion;
    g_table[219] = (void*)f_GetJavaVM;
    g_table[228] = (void*)f_ExceptionCheck;
    g_table[229] = (void*)f_NewDirectByteBuffer;
    g_table[230] = (void*)f_GetDirectBufferAddress;
    for (auto& p : g_vm_table) p = nullptr;
    g_vm_table[5] = (void*)vm_Detach;
    g_vm_table[6] = (void*)vm_GetEnv;
    g_vm_table[7] = (void*)vm_Attach;
}

void dump_metrics(MetricNodeObj* n, const std::string& path, std::string* out) {
    for (auto& kv : n->values) *out += path + ":" + kv.first + "=" + std::to_string(kv.second) + "\n";
    for (auto& c : n->children) dump_metrics(c.second, path + "/" + std::to_string(c.first), out);
}

}  // namespace

// ---- C interface for the Python tests (ctypes) -----------------------------------------------------------------------
extern "C" {

void* mock_new() {
    std::call_once(g_once, init_tables);
    auto* vm = new Jvm;
    vm->vm.functions = g_vm_table;
    vm->vm.jvm = vm;
    return vm;
}
// the JNIEnv of the calling thread (the "Java thread" that calls the natives)
void* mock_env(void* jvm) {
    auto* vm = (Jvm*)jvm;
    if (!(tl_env && tl_env_owner == vm)) {
        tl_env = new_env(vm);
        tl_env_owner = vm;
    }
    if (!tl_env->class_loader) {   // a Spark task thread: has a context class loader and a TaskContext
        tl_env->class_loader = vm->make<Obj>("java/lang/ClassLoader");
        tl_env->thread_context = vm->make<Obj>("org/apache/spark/TaskContext");
    }
    return tl_env;
}
void* mock_wrapper(void* jvm, const uint8_t* task_definition, int64_t len, int metric_depth, int metric_fanout) {
    auto* vm = (Jvm*)jvm;
    auto* w = vm->make<WrapperObj>("org/apache/auron/jni/AuronCallNativeWrapper");
    w->task_definition = vm->make<ByteArrayObj>("[B");
    w->task_definition->bytes.assign(task_definition, task_definition + len);
    // the JVM side builds the MetricNode tree from the Spark plan; the mock builds a full tree of the given depth / fan-out
    std::vector<MetricNodeObj*> level;
    w->metrics = vm->make<MetricNodeObj>("org/apache/auron/metric/MetricNode");
    level.push_back(w->metrics);
    for (int d = 1; d < metric_depth; d++) {
        std::vector<MetricNodeObj*> next;
        for (auto* p : level)
            for (int i = 0; i < metric_fanout; i++) {
                auto* c = vm->make<MetricNodeObj>("org/apache/auron/metric/MetricNode");
                p->children[i] = c;
                next.push_back(c);
            }
        level.swap(next);
    }
    return w;
}
void mock_wrapper_fail_import_after(void* wrapper, int n) { ((WrapperObj*)wrapper)->fail_import_after = n; }
void* mock_class(void* jvm, const char* name) { return ((Jvm*)jvm)->klass(name); }
void mock_put_exporter(void* jvm, const char* resource_id, export_fn fn, void* user) {
    auto* vm = (Jvm*)jvm;
    auto* x = vm->make<ExporterObj>("org/apache/auron/arrowio/AuronArrowFFIExporter");
    x->fn = fn;
    x->user = user;
    std::lock_guard<std::mutex> g(vm->mu);
    vm->resources[resource_id] = x;
}
int mock_exporter_closed(void* jvm, const char* resource_id) {
    auto* vm = (Jvm*)jvm;
    auto* x = dynamic_cast<ExporterObj*>(vm->resources[resource_id]);
    return x ? x->closed : -1;
}
void mock_put_fs_provider(void* jvm, const char* resource_id) {
    auto* vm = (Jvm*)jvm;
    auto* p = vm->make<FsProviderObj>("scala/Function1");
    std::lock_guard<std::mutex> g(vm->mu);
    vm->resources[resource_id] = p;
}
// kind: 0 file segment (path, offset, length) ; 1 direct ByteBuffer ; 2 heap ByteBuffer ; 3 ReadableByteChannel (data, length)
void mock_add_block(void* jvm, const char* resource_id, int kind, const char* path, int64_t offset, int64_t length, const uint8_t* data) {
    auto* vm = (Jvm*)jvm;
    BlocksProviderObj* p;
    {
        std::lock_guard<std::mutex> g(vm->mu);
        p = dynamic_cast<BlocksProviderObj*>(vm->resources.count(resource_id) ? vm->resources[resource_id] : nullptr);
    }
    if (!p) {
        p = vm->make<BlocksProviderObj>("scala/Function0");
        p->it = vm->make<IteratorObj>("scala/collection/Iterator");
        std::lock_guard<std::mutex> g(vm->mu);
        vm->resources[resource_id] = p;
    }
    auto* b = vm->make<BlockObj>("org/apache/spark/sql/execution/auron/shuffle/BlockObject");
    b->kind = kind;
    if (kind == 0) {
        b->path = path;
        b->offset = offset;
        b->length = length;
    } else if (kind == 1 || kind == 2) {
        // both buffers are views with a non-zero position (and, for the heap one, a non-zero arrayOffset) into a larger store
        const int lead = 5, base = kind == 2 ? 3 : 0;
        auto* store = vm->make<ByteArrayObj>("[B");
        store->bytes.assign((size_t)(base + lead + length + 7), 0xEE);
        memcpy(store->bytes.data() + base + lead, data, (size_t)length);
        auto* bb = vm->make<ByteBufferObj>("java/nio/ByteBuffer");
        bb->direct = kind == 1;
        bb->address = kind == 1 ? store->bytes.data() : nullptr;
        bb->array = kind == 2 ? store : nullptr;
        bb->array_offset = base;
        bb->position = lead;
        bb->limit = lead + (int)length;
        b->buffer = bb;
    } else {
        b->channel = vm->make<ChannelObj>("java/nio/channels/ReadableByteChannel");
        b->channel->bytes.assign(data, data + length);
    }
    p->it->blocks.push_back(b);
}
int mock_blocks_closed(void* jvm, const char* resource_id) {
    auto* vm = (Jvm*)jvm;
    auto* p = dynamic_cast<BlocksProviderObj*>(vm->resources[resource_id]);
    int n = 0;
    if (p)
        for (auto* b : p->it->blocks) n += b->closed == 1;
    return n;
}
void mock_set_conf(void* jvm, const char* key, const char* value) {
    auto* vm = (Jvm*)jvm;
    std::lock_guard<std::mutex> g(vm->mu);
    vm->conf[key] = value;
}
void mock_set_task_running(void* jvm, int running) { ((Jvm*)jvm)->task_running = running != 0; }

int mock_take_schema(void* wrapper, void* out) {
    auto* w = (WrapperObj*)wrapper;
    if (!w->have_schema) return 0;
    memcpy(out, &w->schema, sizeof(ArrowSchema));
    w->have_schema = false;
    return 1;
}
int mock_num_batches(void* wrapper) { return (int)((WrapperObj*)wrapper)->batches.size(); }
void mock_take_batch(void* wrapper, int i, void* out) {
    auto* w = (WrapperObj*)wrapper;
    memcpy(out, &w->batches[(size_t)i], sizeof(ArrowArray));
    w->batches[(size_t)i].release = nullptr;
}
// "" when setError was never called; otherwise "message" or "message <- cause class: cause message"
const char* mock_error(void* wrapper) {
    static thread_local std::string s;
    auto* w = (WrapperObj*)wrapper;
    s.clear();
    if (w->error) {
        s = w->error->cls + ": " + w->error->message;
        if (w->error->cause) s += " <- " + w->error->cause->cls + ": " + w->error->cause->message;
    }
    return s.c_str();
}
// exception pending on the calling thread's env: "" or "class: message"; clears it
const char* mock_pending_exception(void* jvm) {
    static thread_local std::string s;
    auto* e = (Env*)mock_env(jvm);
    s.clear();
    if (e->pending) s = e->pending->cls + ": " + e->pending->message;
    e->pending = nullptr;
    return s.c_str();
}
const char* mock_metrics(void* wrapper) {
    static thread_local std::string s;
    s.clear();
    dump_metrics(((WrapperObj*)wrapper)->metrics, "", &s);
    return s.c_str();
}
int64_t mock_counter(void* jvm, const char* name) {
    auto* vm = (Jvm*)jvm;
    std::string n = name;
    std::lock_guard<std::mutex> g(vm->mu);
    if (n == "live_globals") return (int64_t)vm->globals.size();
    if (n == "global_new") return vm->global_new;
    if (n == "attached_threads") return vm->attached;
    if (n == "detached_threads") return vm->detached;
    if (n == "reads_without_context") return vm->reads_without_context;
    if (n == "frames_open") return vm->frames_pushed - vm->frames_popped;
    if (n == "frames_pushed") return vm->frames_pushed;
    if (n == "input_wrappers") {
        int64_t c = 0;
        for (auto& kv : vm->resources) c += kv.first.rfind("__input__:", 0) == 0;
        return c;
    }
    if (n == "input_wrappers_closed") {
        int64_t c = 0;
        for (auto& kv : vm->resources)
            if (kv.first.rfind("__input__:", 0) == 0) c += ((InputWrapperObj*)kv.second)->closed == 1;
        return c;
    }
    if (n == "input_reads") {
        int64_t c = 0;
        for (auto& kv : vm->resources)
            if (kv.first.rfind("__input__:", 0) == 0) c += ((InputWrapperObj*)kv.second)->reads;
        return c;
    }
    return -1;
}
const char* mock_trouble(void* jvm) {
    static thread_local std::string s;
    auto* vm = (Jvm*)jvm;
    std::lock_guard<std::mutex> g(vm->mu);
    s = vm->trouble;
    return s.c_str();
}
void mock_free(void* jvm) {
    auto* vm = (Jvm*)jvm;
    for (auto& kv : vm->resources)
        if (auto* in = dynamic_cast<InputWrapperObj*>(kv.second))
            if (in->fd >= 0) close(in->fd);
    if (tl_env_owner == vm) {
        tl_env = nullptr;
        tl_env_owner = nullptr;
    }
    // Envs of attached worker threads stay referenced by those threads' thread-locals: keep the Jvm itself alive (tests are short-lived)
    vm->heap.clear();
}
}
