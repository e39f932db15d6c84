// tests/emu/nccl.h — TEST INFRASTRUCTURE: the slice of the NCCL API the x-slab host code uses (set-up handshake, the uniform-mass
// and re-balancing all-reduces, the send/recv fallback transport), as an in-process rendezvous between the OS threads that
// play the ranks in the host emulation (tests/emu/emu_nccl.cpp). Found instead of the real <nccl.h> only under
// g++ -DAKUA_HOST_EMU -I tests/emu.
#pragma once
#ifndef AKUA_HOST_EMU
#error "tests/emu/nccl.h is the host emulation shim; compile with -DAKUA_HOST_EMU (tests only)"
#endif
#include <stddef.h>
#include "cuda_runtime.h"

typedef struct EmuNcclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0, ncclInternalError = 3, ncclInvalidArgument = 4 } ncclResult_t;
typedef enum { ncclUint8 = 1, ncclUint32 = 3, ncclUint64 = 5 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclMin = 3 } ncclRedOp_t;

extern "C" {
ncclResult_t ncclGetUniqueId(ncclUniqueId* id);
ncclResult_t ncclCommInitRank(ncclComm_t* comm, int nranks, ncclUniqueId id, int rank);
ncclResult_t ncclCommDestroy(ncclComm_t comm);
ncclResult_t ncclSend(const void* buf, size_t count, ncclDataType_t dt, int peer, ncclComm_t comm, cudaStream_t st);
ncclResult_t ncclRecv(void* buf, size_t count, ncclDataType_t dt, int peer, ncclComm_t comm, cudaStream_t st);
ncclResult_t ncclGroupStart();
ncclResult_t ncclGroupEnd();
ncclResult_t ncclAllReduce(const void* send, void* recv, size_t count, ncclDataType_t dt, ncclRedOp_t op, ncclComm_t comm, cudaStream_t st);
const char* ncclGetErrorString(ncclResult_t r);
}
