// Minimal stand-in for the slice of the TensorFlow 1.x C++ op API that the
// reference's custom-op sources use (REGISTER_OP / OpKernel / OpKernelContext /
// Tensor / shape inference).  TEST INFRASTRUCTURE ONLY: it exists so that the
// reference .cpp files can be compiled *unmodified, where they lie* under
// /root/reference into oracle/_ref/ and driven from tests and the CPU-baseline
// leg of bench.py.  Nothing in the product (rfnet_b200/) includes this.
//
// Written from the call sites in the reference (e.g. pc_distance/tf_nndistance.cpp:3-18,
// 60-122; tf_ops/sampling/tf_sampling.cpp:14-63,95-123; tf_ops/grouping/tf_grouping.cpp:14-110),
// not from TensorFlow's headers.
#ifndef RFNET_ORACLE_TF_STUB_CORE_H_
#define RFNET_ORACLE_TF_STUB_CORE_H_

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

namespace tensorflow {

typedef long long int64;

// ---------------------------------------------------------------- Status
class Status {
 public:
  Status() : ok_(true) {}
  explicit Status(const std::string& msg) : ok_(false), msg_(msg) {}
  static Status OK() { return Status(); }
  bool ok() const { return ok_; }
  const std::string& error_message() const { return msg_; }

 private:
  bool ok_;
  std::string msg_;
};

namespace errors {
template <typename... Args>
inline Status InvalidArgument(Args... args) {
  std::ostringstream os;
  int dummy[] = {0, ((void)(os << args), 0)...};
  (void)dummy;
  return Status(os.str());
}
}  // namespace errors

#define TF_RETURN_IF_ERROR(expr)                 \
  do {                                           \
    ::tensorflow::Status _tf_stub_s = (expr);    \
    if (!_tf_stub_s.ok()) return _tf_stub_s;     \
  } while (0)

// ---------------------------------------------------------------- dtypes
enum DataType { DT_INVALID = 0, DT_FLOAT = 1, DT_INT32 = 3 };
template <typename T> struct DataTypeToEnum;
template <> struct DataTypeToEnum<float> { static const DataType value = DT_FLOAT; };
template <> struct DataTypeToEnum<int> { static const DataType value = DT_INT32; };

static const char* const DEVICE_CPU = "CPU";
static const char* const DEVICE_GPU = "GPU";

// ---------------------------------------------------------------- TensorShape / Tensor
class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<int64> d) : d_(d) {}
  explicit TensorShape(const std::vector<int64>& d) : d_(d) {}
  int dims() const { return (int)d_.size(); }
  int64 dim_size(int i) const { return d_[i]; }
  int64 num_elements() const {
    int64 n = 1;
    for (size_t i = 0; i < d_.size(); i++) n *= d_[i];
    return n;
  }
  bool operator==(const TensorShape& o) const { return d_ == o.d_; }
  bool operator!=(const TensorShape& o) const { return !(d_ == o.d_); }
  const std::vector<int64>& vec() const { return d_; }

 private:
  std::vector<int64> d_;
};

template <typename T>
class FlatView {
 public:
  FlatView(T* p, int64 n) : p_(p), n_(n) {}
  T& operator()(int64 i) const { return p_[i]; }
  int64 size() const { return n_; }
  T* data() const { return p_; }

 private:
  T* p_;
  int64 n_;
};

// A Tensor here never owns caller memory; temporaries own theirs via a shared deleter.
class Tensor {
 public:
  Tensor() : dtype_(DT_INVALID), data_(nullptr) {}
  Tensor(DataType dt, const TensorShape& s, void* data) : dtype_(dt), shape_(s), data_(data) {}
  int dims() const { return shape_.dims(); }
  const TensorShape& shape() const { return shape_; }
  DataType dtype() const { return dtype_; }
  template <typename T>
  FlatView<T> flat() const { return FlatView<T>(static_cast<T*>(data_), shape_.num_elements()); }
  void* raw() const { return data_; }
  void set_owner(std::shared_ptr<void> o) { owner_ = o; }

 private:
  DataType dtype_;
  TensorShape shape_;
  void* data_;
  std::shared_ptr<void> owner_;
};

// ---------------------------------------------------------------- kernels
namespace stub {
struct Attrs { std::map<std::string, int> ints; };
// Device memory hooks; the driver installs cudaMalloc/cudaFree for GPU kernels.
typedef void* (*AllocFn)(size_t);
typedef void (*FreeFn)(void*);
}  // namespace stub

class OpKernelConstruction {
 public:
  explicit OpKernelConstruction(const stub::Attrs* a) : attrs_(a) {}
  Status GetAttr(const char* name, int* out) const {
    std::map<std::string, int>::const_iterator it = attrs_->ints.find(name);
    if (it == attrs_->ints.end()) return Status(std::string("missing attr ") + name);
    *out = it->second;
    return Status::OK();
  }
  void CtxFailure(const Status& s) { if (status_.ok()) status_ = s; }
  const Status& status() const { return status_; }

 private:
  const stub::Attrs* attrs_;
  Status status_;
};

class OpKernelContext {
 public:
  struct Out { void* ptr; int64 capacity_bytes; TensorShape shape; bool allocated; };
  OpKernelContext(const std::vector<Tensor>* in, std::vector<Out>* out, stub::AllocFn a, stub::FreeFn f)
      : in_(in), out_(out), alloc_(a), free_(f) {}
  const Tensor& input(int i) const { return (*in_)[i]; }
  Status allocate_output(int i, const TensorShape& s, Tensor** t) {
    if (i < 0 || i >= (int)out_->size()) return Status("allocate_output: no caller buffer for this output");
    Out& o = (*out_)[i];
    if (s.num_elements() * 4 > o.capacity_bytes) return Status("allocate_output: caller buffer too small");
    o.shape = s;
    o.allocated = true;
    out_tensors_.push_back(std::unique_ptr<Tensor>(new Tensor(DT_FLOAT, s, o.ptr)));
    *t = out_tensors_.back().get();
    return Status::OK();
  }
  Status allocate_temp(DataType dt, const TensorShape& s, Tensor* t) {
    size_t bytes = (size_t)s.num_elements() * 4;
    void* p = alloc_(bytes ? bytes : 4);
    if (!p) return Status("allocate_temp failed");
    stub::FreeFn f = free_;
    std::shared_ptr<void> owner(p, [f](void* q) { f(q); });
    *t = Tensor(dt, s, p);
    t->set_owner(owner);
    return Status::OK();
  }
  void CtxFailure(const Status& s) { if (status_.ok()) status_ = s; }
  const Status& status() const { return status_; }

 private:
  const std::vector<Tensor>* in_;
  std::vector<Out>* out_;
  stub::AllocFn alloc_;
  stub::FreeFn free_;
  std::vector<std::unique_ptr<Tensor> > out_tensors_;
  Status status_;
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext* context) = 0;
};

#define OP_REQUIRES(CTX, EXP, STATUS)      \
  do {                                     \
    if (!(EXP)) {                          \
      (CTX)->CtxFailure((STATUS));         \
      return;                              \
    }                                      \
  } while (0)

#define OP_REQUIRES_OK(CTX, ...)                          \
  do {                                                    \
    ::tensorflow::Status _tf_stub_s(__VA_ARGS__);         \
    if (!_tf_stub_s.ok()) {                               \
      (CTX)->CtxFailure(_tf_stub_s);                      \
      return;                                             \
    }                                                     \
  } while (0)

// ---------------------------------------------------------------- shape inference (parsed, never evaluated)
namespace shape_inference {
struct DimensionHandle { int64 v; DimensionHandle() : v(-1) {} explicit DimensionHandle(int64 x) : v(x) {} };
struct ShapeHandle { std::vector<int64> d; };
struct DimensionOrConstant {
  int64 v;
  DimensionOrConstant(DimensionHandle h) : v(h.v) {}
  DimensionOrConstant(int64 x) : v(x) {}
  DimensionOrConstant(int x) : v(x) {}
};
class InferenceContext {
 public:
  ShapeHandle input(int) { return ShapeHandle(); }
  Status WithRank(ShapeHandle s, int, ShapeHandle* out) { *out = s; return Status::OK(); }
  DimensionHandle Dim(ShapeHandle, int) { return DimensionHandle(); }
  ShapeHandle MakeShape(std::initializer_list<DimensionOrConstant>) { return ShapeHandle(); }
  void set_output(int, ShapeHandle) {}
  Status GetAttr(const char*, int* out) { *out = 0; return Status::OK(); }
};
}  // namespace shape_inference

// ---------------------------------------------------------------- registries
namespace stub {
typedef OpKernel* (*KernelFactory)(OpKernelConstruction*);
struct KernelKey { std::string op, device; bool operator<(const KernelKey& o) const { return op < o.op || (op == o.op && device < o.device); } };
inline std::map<KernelKey, KernelFactory>& kernel_registry() {
  static std::map<KernelKey, KernelFactory> r;
  return r;
}

class OpDefBuilder {
 public:
  explicit OpDefBuilder(const char* name) : name_(name) {}
  OpDefBuilder& Input(const char*) { return *this; }
  OpDefBuilder& Output(const char*) { return *this; }
  OpDefBuilder& Attr(const char*) { return *this; }
  OpDefBuilder& SetShapeFn(Status (*)(shape_inference::InferenceContext*)) { return *this; }
  OpDefBuilder& Doc(const char*) { return *this; }

 private:
  std::string name_;
};

class KernelDefBuilder {
 public:
  explicit KernelDefBuilder(const char* op) { key_.op = op; }
  KernelDefBuilder& Device(const char* d) { key_.device = d; return *this; }
  const KernelKey& key() const { return key_; }

 private:
  KernelKey key_;
};
struct KernelRegistrar {
  KernelRegistrar(const KernelDefBuilder& b, KernelFactory f) { kernel_registry()[b.key()] = f; }
};
}  // namespace stub

inline stub::KernelDefBuilder Name(const char* op) { return stub::KernelDefBuilder(op); }

#define TF_STUB_CAT_(a, b) a##b
#define TF_STUB_CAT(a, b) TF_STUB_CAT_(a, b)
#define REGISTER_OP(name) \
  static ::tensorflow::stub::OpDefBuilder TF_STUB_CAT(tf_stub_op_, __COUNTER__) = ::tensorflow::stub::OpDefBuilder(name)
#define REGISTER_KERNEL_BUILDER(kdef, ...) TF_STUB_REGISTER_KERNEL_(__COUNTER__, kdef, __VA_ARGS__)
#define TF_STUB_REGISTER_KERNEL_(ctr, kdef, ...)                                                       \
  static ::tensorflow::OpKernel* TF_STUB_CAT(tf_stub_make_, ctr)(::tensorflow::OpKernelConstruction* c) { \
    return new __VA_ARGS__(c);                                                                          \
  }                                                                                                     \
  static ::tensorflow::stub::KernelRegistrar TF_STUB_CAT(tf_stub_reg_, ctr)(kdef, TF_STUB_CAT(tf_stub_make_, ctr))

}  // namespace tensorflow
#endif  // RFNET_ORACLE_TF_STUB_CORE_H_
