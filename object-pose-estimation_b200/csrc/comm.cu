// comm.cu — the one collective of the path, inside the library (BASELINE north star: "NCCL over NVLink used only
// to reduce the best-hypothesis score and transform"): a SAC-IA hypothesis pool sharded over the GPUs of one box. Every rank
// evaluates a contiguous share of the SAME pre-drawn decision table on its own device; the ranks then agree on the winner with
// ONE ncclAllReduce(MIN) over a packed 64-bit key — float bits of the error (errors are >= 0, so bit order = numeric order) in the
// high word, hypothesis index in the low word: "first strictly-lower error wins"  becomes a minimum — and the owner of
// the winning hypothesis broadcasts its 4x4 with ONE ncclBroadcast. 8 + 64 bytes per alignment; both on the context's stream.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy the host process already uses — torch's bundled one under torchrun —
// or the system one), so libope_cuda.so has no link-time dependency on it and single-GPU users never load it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "ope_host.cuh"

using namespace ope;

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {std::getenv("OPE_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
    auto sym = [&](const char* s) { return dlsym(api.handle, s); };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.Broadcast && api.GetErrorString;
  });
  return api;
}

__global__ void pack_key_kernel(const ope_reg_result* __restrict__ local, int have, unsigned long long* __restrict__ key) {
  // empty shard: the largest key, never the minimum unless every shard is empty
  unsigned long long k = ~0ull;
  if (have && local->best_iteration >= 0) k = ((unsigned long long)__float_as_uint((float)local->best_error) << 32) | (unsigned)local->best_iteration;
  *key = k;
}

}  // namespace

struct ope_comm {
  ope_ctx* ctx = nullptr;
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0;
  unsigned long long* d_key = nullptr;   // [0] local / reduced key
  float* d_T = nullptr;                  // 16 floats
  ope_reg_result* d_local = nullptr;
};

#define OPE_NCCL_TRY(ctx, expr)                                                                                        \
  do {                                                                                                                 \
    ncclResult_t r__ = (expr);                                                                                         \
    if (r__ != ncclSuccess) return fail((ctx), OPE_ERR_CUDA, "%s failed: %s", #expr, nccl().GetErrorString(r__));      \
  } while (0)

extern "C" {

int ope_comm_unique_id(void* out, size_t bytes) {
  if (!out || bytes < sizeof(ncclUniqueId)) return OPE_ERR_INVALID;
  if (!nccl().ok) return OPE_ERR_UNSUPPORTED;
  ncclUniqueId id;
  if (nccl().GetUniqueId(&id) != ncclSuccess) return OPE_ERR_CUDA;
  std::memcpy(out, &id, sizeof(id));
  return OPE_OK;
}

int ope_comm_nccl_version(void) {
  int v = 0;
  if (!nccl().ok || !nccl().GetVersion || nccl().GetVersion(&v) != ncclSuccess) return 0;
  return v;
}

int ope_comm_create(ope_ctx* ctx, const void* unique_id, size_t bytes, int world, int rank, ope_comm** out) {
  OPE_ENTER(ctx);
  if (!ctx || !out || world < 1 || rank < 0 || rank >= world) return OPE_ERR_INVALID;
  *out = nullptr;
  ope_comm* c = new ope_comm();
  c->ctx = ctx; c->world = world; c->rank = rank;
  int rc = dalloc(ctx, &c->d_key, 1);
  if (rc == OPE_OK) rc = dalloc(ctx, &c->d_T, 16);
  if (rc == OPE_OK) rc = dalloc(ctx, &c->d_local, 1);
  if (rc == OPE_OK && world > 1) {
    if (!nccl().ok) rc = fail(ctx, OPE_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded (set OPE_NCCL_LIB)");
    else if (!unique_id || bytes < sizeof(ncclUniqueId)) rc = fail(ctx, OPE_ERR_INVALID, "ope_comm_create needs the %zu-byte NCCL unique id of rank 0", sizeof(ncclUniqueId));
    else {
      ncclUniqueId id;
      std::memcpy(&id, unique_id, sizeof(id));
      const ncclResult_t r = nccl().CommInitRank(&c->comm, world, id, rank);
      if (r != ncclSuccess) rc = fail(ctx, OPE_ERR_CUDA, "ncclCommInitRank failed: %s", nccl().GetErrorString(r));
    }
  }
  if (rc != OPE_OK) { dfree(ctx, c->d_key); dfree(ctx, c->d_T); dfree(ctx, c->d_local); delete c; return rc; }
  *out = c;
  return OPE_OK;
}

void ope_comm_destroy(ope_comm* c) {
  if (!c) return;
  OPE_ENTER(c->ctx);
  stream_sync(c->ctx);
  if (c->comm) nccl().CommDestroy(c->comm);
  dfree(c->ctx, c->d_key); dfree(c->ctx, c->d_T); dfree(c->ctx, c->d_local);
  delete c;
}

int ope_comm_rank(const ope_comm* c) { return c ? c->rank : 0; }
int ope_comm_world(const ope_comm* c) { return c ? c->world : 1; }

// SampleConsensusInitialAlignment::align over a pool sharded across the communicator's ranks. Every rank passes the SAME clouds,
// features and pre-drawn table (ope_sacia_draw after a common srand, or broadcast by the caller); rank r evaluates hypotheses
// [H r / world, H (r + 1) / world). On return every rank holds the pool's winner — exactly the single-GPU result.
int ope_sacia_align_sharded(ope_ctx* ctx, ope_comm* comm, const ope_cloud* src, const float* fsrc, const ope_cloud* tgt, const float* ftgt,
                            const ope_sacia_params* prm, const ope_rng_table* table, ope_reg_result* res) {
  OPE_ENTER(ctx);
  if (!ctx || !comm || !src || !tgt || !fsrc || !ftgt || !prm || !res) return OPE_ERR_INVALID;
  if (!table) return fail(ctx, OPE_ERR_INVALID, "a sharded pool needs the pre-drawn decision table (every rank must evaluate the same pool)");
  if (comm->ctx != ctx) return fail(ctx, OPE_ERR_INVALID, "the communicator belongs to another context");
  const int H = prm->max_iterations, world = comm->world, rank = comm->rank;
  ope_sacia_params p = *prm;
  p.hypothesis_begin = (int)((long long)H * rank / world);
  p.hypothesis_end = (int)((long long)H * (rank + 1) / world);
  ope_reg_result local;
  std::memset(&local, 0, sizeof(local));
  local.best_iteration = -1;
  const int have = p.hypothesis_end > p.hypothesis_begin ? 1 : 0;
  if (have) OPE_TRY(ope_sacia_align(ctx, src, fsrc, tgt, ftgt, &p, table, &local, nullptr));
  if (world == 1) { *res = local; return OPE_OK; }
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(comm->d_local, &local, sizeof(local), cudaMemcpyHostToDevice, ctx->stream));
  OPE_CUDA_TRY(ctx, cudaMemcpyAsync(comm->d_T, local.T, 16 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  pack_key_kernel<<<1, 1, 0, ctx->stream>>>(comm->d_local, have, comm->d_key);
  OPE_TRY(check_launch(ctx, "pack_key_kernel"));
  OPE_NCCL_TRY(ctx, nccl().AllReduce(comm->d_key, comm->d_key, 1, ncclUint64, ncclMin, comm->comm, ctx->stream));
  void* h;
  OPE_TRY(read_back(ctx, comm->d_key, sizeof(unsigned long long), &h));
  const unsigned long long best = *(const unsigned long long*)h;
  std::memset(res, 0, sizeof(*res));
  for (int i = 0; i < 16; ++i) res->T[i] = (i % 5 == 0) ? 1.0f : 0.0f;
  res->best_iteration = -1; res->iterations = H;
  if (best == ~0ull) return OPE_OK;   // no rank had a hypothesis
  const int winner = (int)(best & 0xffffffffull);
  int owner = 0;
  for (int r = 0; r < world; ++r)
    if (winner >= (int)((long long)H * r / world) && winner < (int)((long long)H * (r + 1) / world)) owner = r;
  OPE_NCCL_TRY(ctx, nccl().Broadcast(comm->d_T, comm->d_T, 16, ncclFloat32, owner, comm->comm, ctx->stream));
  OPE_TRY(read_back(ctx, comm->d_T, 16 * sizeof(float), &h));
  std::memcpy(res->T, h, 16 * sizeof(float));
  const unsigned bits = (unsigned)(best >> 32);
  float e;
  std::memcpy(&e, &bits, 4);
  res->best_error = (double)e; res->best_iteration = winner; res->converged = 1;
  res->last_mse = local.last_mse;
  return OPE_OK;
}

}  // extern "C"
