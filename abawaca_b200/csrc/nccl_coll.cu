// abw_collectives backed directly by NCCL on the context stream (no host synchronisation, no interpreter in the loop).
// libnccl is not linked: it is resolved at run time (the process that drives several GPUs has it loaded already, e.g. through torch.distributed),
// so that single-GPU users of libabawaca_b200.so need no NCCL at all.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>          // types and enumerators only

namespace {

struct NcclApi {
	void* lib = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl_api(std::string* why)
{
	static NcclApi api;
	static bool tried = false, ok = false;
	static std::string err;
	if(!tried) {
		tried = true;
		const char* names[] = {getenv("ABW_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
		for(const char* n : names) {
			if(!n)
				continue;
			api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
			if(api.lib)
				break;
		}
		if(!api.lib)
			err = "libnccl.so.2 not found (set ABW_NCCL_LIB)";
		else {
			api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
			api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
			api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
			api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
			api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
			api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
			ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.AllReduce;
			if(!ok)
				err = "libnccl lacks an expected symbol";
		}
	}
	if(!ok && why)
		*why = err;
	return ok? &api : nullptr;
}

struct NcclUser {
	NcclApi* api;
	ncclComm_t comm;
	abw_ctx* ctx;
};

int cb_allgather(void* user, const void* d_send, void* d_recv, size_t bytes_per_rank)
{
	NcclUser* u = (NcclUser*)user;
	return u->api->AllGather(d_send, d_recv, bytes_per_rank, ncclUint8, u->comm, u->ctx->stream) == ncclSuccess? 0 : 1;     // stream ordered: later work on the context waits for it
}

int cb_allreduce(void* user, void* d_buf, size_t count)
{
	NcclUser* u = (NcclUser*)user;
	return u->api->AllReduce(d_buf, d_buf, count, ncclInt64, ncclSum, u->comm, u->ctx->stream) == ncclSuccess? 0 : 1;
}

}  // namespace

extern "C" {

int abw_nccl_unique_id(abw_ctx* ctx, void* id128)
{
	if(!ctx || !id128)
		return abw_fail(ctx, ABW_ERR_ARG, "abw_nccl_unique_id: null argument");
	std::string why;
	NcclApi* api = nccl_api(&why);
	if(!api)
		return abw_fail(ctx, ABW_ERR_UNSUPPORTED, "abw_nccl_unique_id: " + why);
	static_assert(sizeof(ncclUniqueId) == ABW_NCCL_ID_BYTES, "ncclUniqueId size");
	ncclUniqueId id;
	if(api->GetUniqueId(&id) != ncclSuccess)
		return abw_fail(ctx, ABW_ERR_CUDA, "ncclGetUniqueId failed");
	memcpy(id128, &id, sizeof(id));
	return ABW_OK;
}

int abw_nccl_collectives_create(abw_ctx* ctx, const void* id128, int rank, int world, abw_collectives* out)
{
	if(!ctx || !id128 || !out || world < 1 || rank < 0 || rank >= world)
		return abw_fail(ctx, ABW_ERR_ARG, "abw_nccl_collectives_create: bad argument");
	ABW_ENTER(ctx);
	std::string why;
	NcclApi* api = nccl_api(&why);
	if(!api)
		return abw_fail(ctx, ABW_ERR_UNSUPPORTED, "abw_nccl_collectives_create: " + why);
	ncclUniqueId id;
	memcpy(&id, id128, sizeof(id));
	ncclComm_t comm;
	ncclResult_t r = api->CommInitRank(&comm, world, id, rank);
	if(r != ncclSuccess)
		return abw_fail(ctx, ABW_ERR_CUDA, std::string("ncclCommInitRank failed: ") + (api->GetErrorString? api->GetErrorString(r) : "?"));
	NcclUser* u = new NcclUser{api, comm, ctx};
	out->allgather = cb_allgather;
	out->allreduce_sum_i64 = cb_allreduce;
	out->user = u;
	out->rank = rank;
	out->world = world;
	out->stream_ordered = 1;
	return ABW_OK;
}

void abw_nccl_collectives_destroy(abw_collectives* c)
{
	if(!c || !c->user)
		return;
	NcclUser* u = (NcclUser*)c->user;
	cudaStreamSynchronize(u->ctx->stream);
	u->api->CommDestroy(u->comm);
	delete u;
	c->user = nullptr;
}

}  // extern "C"
