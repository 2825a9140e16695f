// Driver for the reference build under oracle/_ref/ (TEST INFRASTRUCTURE ONLY).
//
// The reference's op sources are compiled unmodified, where they lie under /root/reference, against
// oracle/tf_stub (see Makefile).  Their REGISTER_KERNEL_BUILDER statics fill tensorflow::stub::kernel_registry();
// this file exposes one C entry point that instantiates a registered kernel by (op name, device) and runs
// its Compute() over caller-owned buffers -- the same thing the TF executor does at
// e.g. pc_distance/tf_nndistance.cpp:60-81 (CPU) or :169-205 (GPU).
//
// CPU kernels take host pointers.  GPU kernels (only present in libref_gpu.so, which also links the reference
// .cu files compiled for sm_100a) take device pointers and launch on the legacy default stream, as the
// reference does (no stream argument anywhere, e.g. tf_ops/CD/tf_nndistance_g.cu:128-129).
#include <cstdio>
#include <string>
#include <vector>

#include "tensorflow/core/framework/op_kernel.h"

#ifdef REF_WITH_CUDA
#include <cuda_runtime.h>
static void* dev_alloc(size_t n) { void* p = nullptr; return cudaMalloc(&p, n) == cudaSuccess ? p : nullptr; }
static void dev_free(void* p) { cudaFree(p); }
#endif
static void* host_alloc(size_t n) { return malloc(n); }
static void host_free(void* p) { free(p); }

static thread_local std::string g_last_error;

extern "C" {

const char* ref_last_error() { return g_last_error.c_str(); }

// Number of kernels registered (sanity check for the loader test).
int ref_num_kernels() { return (int)tensorflow::stub::kernel_registry().size(); }

int ref_has_kernel(const char* op, const char* device) {
  tensorflow::stub::KernelKey k; k.op = op; k.device = device;
  return tensorflow::stub::kernel_registry().count(k) ? 1 : 0;
}

// inputs:  in_ptr[i], in_dtype[i] (1=float32, 3=int32), in_rank[i], in_dims[4*i .. 4*i+rank)
// attrs:   attr_name[j] / attr_val[j] (ints only; that is all the reference ops use)
// outputs: out_ptr[i] with out_cap_bytes[i]; on return out_rank[i], out_dims[4*i..] hold the shape the kernel asked for
// returns 0 on success, 1 on an OP_REQUIRES failure (message via ref_last_error), 2 if no such kernel
int ref_run_op(const char* op, const char* device,
               int n_in, void* const* in_ptr, const int* in_dtype, const int* in_rank, const long long* in_dims,
               int n_attr, const char* const* attr_name, const int* attr_val,
               int n_out, void* const* out_ptr, const long long* out_cap_bytes, int* out_rank, long long* out_dims) {
  using namespace tensorflow;
  g_last_error.clear();
  stub::KernelKey key; key.op = op; key.device = device;
  std::map<stub::KernelKey, stub::KernelFactory>::iterator it = stub::kernel_registry().find(key);
  if (it == stub::kernel_registry().end()) { g_last_error = std::string("no kernel ") + op + "/" + device; return 2; }

  stub::Attrs attrs;
  for (int j = 0; j < n_attr; j++) attrs.ints[attr_name[j]] = attr_val[j];
  OpKernelConstruction cons(&attrs);
  std::unique_ptr<OpKernel> kernel(it->second(&cons));
  if (!cons.status().ok()) { g_last_error = cons.status().error_message(); return 1; }

  std::vector<Tensor> inputs;
  for (int i = 0; i < n_in; i++) {
    std::vector<int64> d(in_dims + 4 * i, in_dims + 4 * i + in_rank[i]);
    inputs.push_back(Tensor((DataType)in_dtype[i], TensorShape(d), in_ptr[i]));
  }
  std::vector<OpKernelContext::Out> outs(n_out);
  for (int i = 0; i < n_out; i++) { outs[i].ptr = out_ptr[i]; outs[i].capacity_bytes = out_cap_bytes[i]; outs[i].allocated = false; }

  stub::AllocFn a = host_alloc; stub::FreeFn f = host_free;
#ifdef REF_WITH_CUDA
  if (std::string(device) == "GPU") { a = dev_alloc; f = dev_free; }
#endif
  {
    OpKernelContext ctx(&inputs, &outs, a, f);
    kernel->Compute(&ctx);
#ifdef REF_WITH_CUDA
    // Temps are freed when ctx dies; cudaFree synchronises, so kernels using them have finished by then.
    if (std::string(device) == "GPU") {
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) { g_last_error = std::string("CUDA: ") + cudaGetErrorString(e); return 1; }
    }
#endif
    if (!ctx.status().ok()) { g_last_error = ctx.status().error_message(); return 1; }
  }
  for (int i = 0; i < n_out; i++) {
    out_rank[i] = outs[i].allocated ? outs[i].shape.dims() : -1;
    for (int k = 0; k < 4; k++) out_dims[4 * i + k] = (outs[i].allocated && k < outs[i].shape.dims()) ? outs[i].shape.dim_size(k) : 0;
  }
  return 0;
}

}  // extern "C"

#ifndef REF_WITH_CUDA
// The CPU-only library still contains the reference's *GpuOp classes (same translation units), which reference
// the CUDA launchers.  They are never registered for use here (tests only run DEVICE_CPU kernels from this
// library), but the symbols must resolve for dlopen(RTLD_NOW).
#define REF_NO_GPU(sig) void sig { fprintf(stderr, "oracle/_ref: GPU launcher called in the CPU-only reference build\n"); abort(); }
REF_NO_GPU(NmDistanceKernelLauncher(int, int, const float*, int, const float*, float*, int*, float*, int*))
REF_NO_GPU(NmDistanceGradKernelLauncher(int, int, const float*, int, const float*, const float*, const int*, const float*, const int*, float*, float*))
REF_NO_GPU(approxmatchLauncher(int, int, int, const float*, const float*, float*, float*))
REF_NO_GPU(matchcostLauncher(int, int, int, const float*, const float*, const float*, float*))
REF_NO_GPU(matchcostgradLauncher(int, int, int, const float*, const float*, const float*, float*, float*))
#endif
