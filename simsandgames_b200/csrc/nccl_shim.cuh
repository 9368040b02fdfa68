// NCCL bound at run time.  Included by the translation units that talk to NCCL (sor_shard.cu, flip.cu).
#pragma once
#include <nccl.h>   // types only
#include <dlfcn.h>
#include "common.cuh"

namespace fg {

// NCCL through dlopen/dlsym.  Linking libnccl at build time would pull the system NCCL into every process that
// loads this library; a host program that later loads a newer NCCL of the same SONAME (PyTorch bundles one) then
// fails to resolve its symbols.  Binding lazily means: whichever libnccl.so.2 the process already holds is used,
// and it is only loaded at all when a sharded object is created.
struct NcclApi {
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
	bool ok = false;
};
NcclApi& nccl_api();
int load_nccl();
#define g_nccl nccl_api()
#ifdef FG_NCCL_SHIM_IMPL
NcclApi& nccl_api() { static NcclApi api; return api; }
int load_nccl() {
	if (g_nccl.ok) return FLIPGPU_OK;
	void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
	if (!lib) { fg::set_error("NCCL is not available: %s", dlerror()); return FLIPGPU_ENCCL; }
#define FG_SYM(field, name) g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(lib, name)); if (!g_nccl.field) { fg::set_error("NCCL symbol %s missing", name); return FLIPGPU_ENCCL; }
	FG_SYM(GetUniqueId, "ncclGetUniqueId") FG_SYM(CommInitRank, "ncclCommInitRank") FG_SYM(CommDestroy, "ncclCommDestroy")
	FG_SYM(GroupStart, "ncclGroupStart") FG_SYM(GroupEnd, "ncclGroupEnd") FG_SYM(Send, "ncclSend") FG_SYM(Recv, "ncclRecv")
	FG_SYM(GetErrorString, "ncclGetErrorString") FG_SYM(AllReduce, "ncclAllReduce")
#undef FG_SYM
	g_nccl.ok = true;
	return FLIPGPU_OK;
}
#endif

#define FG_NCCL(call)                                                                          \
	do {                                                                                       \
		ncclResult_t r_ = (call);                                                              \
		if (r_ != ncclSuccess) {                                                               \
			fg::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, fg::g_nccl.GetErrorString(r_)); \
			return FLIPGPU_ENCCL;                                                              \
		}                                                                                      \
	} while (0)

}  // namespace fg
