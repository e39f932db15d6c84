// tests/emu/cuda_host_shim.h — TEST INFRASTRUCTURE: the slice of the CUDA runtime HOST API that akuaengine_b200/csrc/pbf_solver.cu
// uses, implemented on plain host memory, so that the product's host solver itself (not a re-statement of it) can be compiled
// by g++ -DAKUA_HOST_EMU against the SIMT emulator and driven through the real C ABI on the CPU (tests/test_emu_slab.py).
// Included at the end of tests/emu/cuda_runtime.h; never seen by nvcc.
//
// Model: every call is synchronous (streams and events are inert), device memory is malloc'd host memory, an "IPC handle" is
// the pointer itself (all emulated ranks are threads of one process), stream capture is refused (the solver then steps
// eagerly), and the GL-interop entry points fail. NCCL is replaced by tests/emu/nccl.h (in-process rendezvous).
#pragma once

typedef void* cudaEvent_t;
typedef void* cudaGraph_t;
typedef void* cudaGraphExec_t;
typedef struct cudaGraphicsResource* cudaGraphicsResource_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1 };
constexpr unsigned cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1;
constexpr cudaError_t cudaErrorNotSupported = 801, cudaErrorMemoryAllocation = 2;
struct cudaDeviceProp { int major = 10, minor = 0; char name[64] = "emulated sm_100a"; };
struct cudaIpcMemHandle_t { char reserved[64]; };

inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorNotSupported ? "not supported by the host emulation" : "emulated CUDA error"); }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 8; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { *p = cudaDeviceProp{}; return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return cudaSuccess; }

inline cudaError_t cudaMalloc(void** p, size_t bytes) { *p = std::calloc(bytes ? bytes : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t bytes) { return cudaMalloc(reinterpret_cast<void**>(p), bytes); }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void** p, size_t bytes) { return cudaMalloc(p, bytes); }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }

inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorNotSupported; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = nullptr; return cudaErrorNotSupported; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t, unsigned long long) { *e = nullptr; return cudaErrorNotSupported; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }

inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }

inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { std::memset(h, 0, sizeof(*h)); std::memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { std::memcpy(p, h.reserved, sizeof(*p)); return cudaSuccess; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }

inline cudaError_t cudaGraphicsMapResources(int, cudaGraphicsResource_t*, cudaStream_t) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphicsUnmapResources(int, cudaGraphicsResource_t*, cudaStream_t) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphicsResourceGetMappedPointer(void**, size_t*, cudaGraphicsResource_t) { return cudaErrorNotSupported; }
