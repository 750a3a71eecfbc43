// sp.cu — Ulysses sequence parallelism for the joint attention (new in this build: the reference is single-device,
// SURVEY.md §2a / §8e).
//
// Tokens of each stream are sharded contiguously over P ranks (rank r holds txt rows [r St/P, (r+1) St/P) and img rows
// [r Si/P, (r+1) Si/P)); every linear / elementwise kernel is local. Around the attention of each block
//   Q,K,V : [S/P tokens, H heads]  ->  [S tokens, H/P heads]        (exchange 1)
//   O     : [S tokens, H/P heads]  ->  [S/P tokens, H heads]        (exchange 2)
// QK-RMSNorm and RoPE are per (token, head) and were already applied by the QKV GEMM epilogue; softmax is invariant to
// the order of the keys and there is no mask, so gathered rows simply stay in rank-major order.
//
// Two transports:
//   mode 0  NCCL: grouped ncclSend / ncclRecv (an all-to-all) on the context stream. The QKV epilogue has already
//           written the send buffer in [dest][token][q|k|v][H/P*128] order, so every message is one contiguous slab.
//   mode 1  peer memory: the QKV GEMM epilogue and the attention epilogue store straight into the destination rank's
//           buffers over NVLink (cudaIpc mappings); the only extra kernel is a flag barrier. NCCL is then used for
//           set-up (handle exchange) and the final [S_img, 128] all-gather only.
// NCCL is resolved with dlopen at first use so that libflux2b.so has no hard dependency on it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <vector>

#include "ctx.h"

namespace f2b {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
static NcclApi g_nccl;

static bool nccl_load() {
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return;
#define F2B_SYM(field, name) *reinterpret_cast<void**>(&g_nccl.field) = dlsym(g_nccl.lib, name)
    F2B_SYM(GetUniqueId, "ncclGetUniqueId");
    F2B_SYM(CommInitRank, "ncclCommInitRank");
    F2B_SYM(CommDestroy, "ncclCommDestroy");
    F2B_SYM(GroupStart, "ncclGroupStart");
    F2B_SYM(GroupEnd, "ncclGroupEnd");
    F2B_SYM(Send, "ncclSend");
    F2B_SYM(Recv, "ncclRecv");
    F2B_SYM(AllGather, "ncclAllGather");
    F2B_SYM(GetErrorString, "ncclGetErrorString");
#undef F2B_SYM
    g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.GroupStart && g_nccl.GroupEnd &&
                g_nccl.Send && g_nccl.Recv && g_nccl.AllGather && g_nccl.GetErrorString;
  });
  return g_nccl.ok;
}

#define F2B_NCCL(expr)                                                                                         \
  do {                                                                                                         \
    ncclResult_t _r = (expr);                                                                                  \
    if (_r != ncclSuccess) return fail(FLUX2B_ERR_CUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(_r)); \
  } while (0)

int sp_fork(flux2b_ctx* c) {
  SpState& sp = c->sp;
  if (!sp.comm_stream) {
    F2B_CUDA(cudaStreamCreateWithFlags(&sp.comm_stream, cudaStreamNonBlocking));
    F2B_CUDA(cudaEventCreateWithFlags(&sp.ev_main, cudaEventDisableTiming));
    F2B_CUDA(cudaEventCreateWithFlags(&sp.ev_comm, cudaEventDisableTiming));
  }
  F2B_CUDA(cudaEventRecord(sp.ev_main, c->stream));
  F2B_CUDA(cudaStreamWaitEvent(sp.comm_stream, sp.ev_main, 0));
  return 0;
}
int sp_join(flux2b_ctx* c) {
  SpState& sp = c->sp;
  F2B_CUDA(cudaEventRecord(sp.ev_comm, sp.comm_stream));
  F2B_CUDA(cudaStreamWaitEvent(c->stream, sp.ev_comm, 0));
  return 0;
}

int sp_all_to_all(flux2b_ctx* c, const void* send, void* recv, size_t chunk_elems16, cudaStream_t stream) {
  if (!stream) stream = c->stream;
  ncclComm_t comm = reinterpret_cast<ncclComm_t>(c->sp.comm);
  const int P = c->sp.world;
  const uint16_t* s = reinterpret_cast<const uint16_t*>(send);
  uint16_t* r = reinterpret_cast<uint16_t*>(recv);
  // (profiler events are recorded on the context stream, so an exchange on the side stream only counts launches / bytes)
  c->launches++;
  c->prof[FLUX2B_PROF_COMM].launches++;
  c->prof[FLUX2B_PROF_COMM].bytes += 2.0 * chunk_elems16 * 2 * (P - 1);
  F2B_NCCL(g_nccl.GroupStart());
  for (int peer = 0; peer < P; ++peer) {
    F2B_NCCL(g_nccl.Send(s + (size_t)peer * chunk_elems16, chunk_elems16, ncclBfloat16, peer, comm, stream));
    F2B_NCCL(g_nccl.Recv(r + (size_t)peer * chunk_elems16, chunk_elems16, ncclBfloat16, peer, comm, stream));
  }
  F2B_NCCL(g_nccl.GroupEnd());
  return 0;
}

int sp_all_gather_f32(flux2b_ctx* c, float* buf, size_t elems_per_rank) {
  ncclComm_t comm = reinterpret_cast<ncclComm_t>(c->sp.comm);
  ProfScope ps(c, FLUX2B_PROF_COMM, 0, 4.0 * elems_per_rank * (c->sp.world - 1));
  F2B_NCCL(g_nccl.AllGather(buf + (size_t)c->sp.rank * elems_per_rank, buf, elems_per_rank, ncclFloat, comm, c->stream));
  return 0;
}

static void sp_unmap(flux2b_ctx* c) {
  for (int i = 0; i < 8; ++i) {
    if (i != c->sp.rank) {
      if (c->sp.gather_peer[i]) cudaIpcCloseMemHandle(c->sp.gather_peer[i]);
      if (c->sp.cat_peer[i]) cudaIpcCloseMemHandle(c->sp.cat_peer[i]);
      if (c->sp.flag_peer[i]) cudaIpcCloseMemHandle(c->sp.flag_peer[i]);
    }
    c->sp.gather_peer[i] = c->sp.cat_peer[i] = nullptr;
    c->sp.flag_peer[i] = nullptr;
  }
  c->sp.gather_exported = c->sp.cat_exported = c->sp.flag_exported = nullptr;
}

// Exchange cudaIpc handles of the three buffers peers write into. Collective: every rank reaches the same decision
// because all ranks run the same shapes in the same order (the buffers grow at the same calls).
int sp_map_peers(flux2b_ctx* c) {
  SpState& sp = c->sp;
  if (!c->ws_sp_flags.p) {
    F2B_CUDA(c->ws_sp_flags.alloc(256));
    F2B_CUDA(cudaMemsetAsync(c->ws_sp_flags.p, 0, 256, c->stream));
  }
  if (sp.gather_exported == c->ws_sp_gather.p && sp.cat_exported == c->ws_cat.p && sp.flag_exported == c->ws_sp_flags.p) return 0;
  F2B_CUDA(cudaStreamSynchronize(c->stream));
  sp_unmap(c);
  struct Handles { cudaIpcMemHandle_t g, x, f; };
  Handles mine;
  F2B_CUDA(cudaIpcGetMemHandle(&mine.g, c->ws_sp_gather.p));
  F2B_CUDA(cudaIpcGetMemHandle(&mine.x, c->ws_cat.p));
  F2B_CUDA(cudaIpcGetMemHandle(&mine.f, c->ws_sp_flags.p));
  DevBuf dev;
  F2B_CUDA(dev.alloc(sizeof(Handles) * sp.world));
  F2B_CUDA(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(dev.p) + sizeof(Handles) * sp.rank, &mine, sizeof(Handles), cudaMemcpyHostToDevice, c->stream));
  F2B_NCCL(g_nccl.AllGather(reinterpret_cast<uint8_t*>(dev.p) + sizeof(Handles) * sp.rank, dev.p, sizeof(Handles), ncclUint8,
                            reinterpret_cast<ncclComm_t>(sp.comm), c->stream));
  std::vector<Handles> all(sp.world);
  F2B_CUDA(cudaMemcpyAsync(all.data(), dev.p, sizeof(Handles) * sp.world, cudaMemcpyDeviceToHost, c->stream));
  F2B_CUDA(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < sp.world; ++i) {
    if (i == sp.rank) {
      sp.gather_peer[i] = c->ws_sp_gather.p; sp.cat_peer[i] = c->ws_cat.p; sp.flag_peer[i] = c->ws_sp_flags.as<uint32_t>();
      continue;
    }
    void* q = nullptr;
    F2B_CUDA(cudaIpcOpenMemHandle(&q, all[i].g, cudaIpcMemLazyEnablePeerAccess)); sp.gather_peer[i] = q;
    F2B_CUDA(cudaIpcOpenMemHandle(&q, all[i].x, cudaIpcMemLazyEnablePeerAccess)); sp.cat_peer[i] = q;
    F2B_CUDA(cudaIpcOpenMemHandle(&q, all[i].f, cudaIpcMemLazyEnablePeerAccess)); sp.flag_peer[i] = reinterpret_cast<uint32_t*>(q);
  }
  sp.gather_exported = c->ws_sp_gather.p; sp.cat_exported = c->ws_cat.p; sp.flag_exported = c->ws_sp_flags.p;
  // nobody may write into the new mappings before everyone has opened them; flags restart from a clean epoch
  F2B_CUDA(cudaMemsetAsync(c->ws_sp_flags.p, 0, 256, c->stream));
  sp.epoch = 0;
  F2B_NCCL(g_nccl.AllGather(reinterpret_cast<uint8_t*>(dev.p) + sp.rank, dev.p, 1, ncclUint8, reinterpret_cast<ncclComm_t>(sp.comm), c->stream));
  F2B_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// Flag barrier over peer memory: rank r stores `epoch` into slot r of every rank's flag array (system-scope release
// after a system fence: the peer stores of the preceding kernels are ordered before it), then waits until all slots of
// its own array reached `epoch`.
struct PeerFlags { uint32_t* p[8]; };
__global__ void sp_barrier_kernel(PeerFlags peers, uint32_t* mine, int rank, int world, uint32_t epoch) {
  const int t = threadIdx.x;
  if (t < world) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peers.p[t] + rank), "r"(epoch) : "memory");
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine + t) : "memory");
    } while ((int32_t)(v - epoch) < 0);
  }
}
int sp_barrier(flux2b_ctx* c) {
  SpState& sp = c->sp;
  PeerFlags pf{};
  for (int i = 0; i < sp.world; ++i) pf.p[i] = sp.flag_peer[i];
  ++sp.epoch;
  ProfScope ps(c, FLUX2B_PROF_COMM, 0, 0);
  sp_barrier_kernel<<<1, 32, 0, c->stream>>>(pf, c->ws_sp_flags.as<uint32_t>(), sp.rank, sp.world, sp.epoch);
  F2B_CUDA(cudaGetLastError());
  return 0;
}

void sp_destroy(flux2b_ctx* c) {
  sp_unmap(c);
  if (c->sp.comm_stream) {
    cudaStreamSynchronize(c->sp.comm_stream);
    cudaEventDestroy(c->sp.ev_main); cudaEventDestroy(c->sp.ev_comm);
    cudaStreamDestroy(c->sp.comm_stream);
    c->sp.comm_stream = nullptr; c->sp.ev_main = c->sp.ev_comm = nullptr;
  }
  if (c->sp.comm && g_nccl.ok) g_nccl.CommDestroy(reinterpret_cast<ncclComm_t>(c->sp.comm));
  c->sp.comm = nullptr;
  c->sp.world = 1;
  c->sp.rank = 0;
}

}  // namespace f2b

using namespace f2b;

extern "C" {

int flux2b_sp_unique_id(void* id128) {
  if (!id128) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "null id buffer");
  if (!nccl_load()) return fail(FLUX2B_ERR_NO_DEVICE, "libnccl.so.2 could not be loaded");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  F2B_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int flux2b_sp_layout(int world, int rank, int S_txt, int S_img, int num_heads, flux2b_sp_layout_t* out) {
  if (!out || world < 1 || world > 8 || rank < 0 || rank >= world) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "bad rank / world (1..8 ranks of one node)");
  if (S_txt < 1 || S_img < 1 || num_heads < 1) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "empty sequence");
  if (S_txt % world || S_img % world) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "S_txt and S_img must be divisible by the sequence-parallel world size");
  if (num_heads % world) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "num_attention_heads must be divisible by the world size");
  out->txt_rows = S_txt / world; out->txt_row0 = rank * out->txt_rows;
  out->img_rows = S_img / world; out->img_row0 = rank * out->img_rows;
  out->local_rows = out->txt_rows + out->img_rows;
  out->heads_per_rank = num_heads / world;
  out->qkv_chunk_elems = (int64_t)out->local_rows * 3 * out->heads_per_rank * 128;
  out->o_chunk_elems = (int64_t)out->local_rows * out->heads_per_rank * 128;
  return 0;
}

int flux2b_sp_init(flux2b_ctx* c, const void* id128, int rank, int world) {
  if (!c || !id128) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "null argument");
  if (world < 1 || world > 8 || rank < 0 || rank >= world) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "bad rank / world (1..8 ranks of one node)");
  if (cudaSetDevice(c->device) != cudaSuccess) return fail(FLUX2B_ERR_NO_DEVICE, "cudaSetDevice failed");
  if (c->has_dit && (c->dit.num_attention_heads % world)) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "num_attention_heads must be divisible by the world size");
  if (!nccl_load()) return fail(FLUX2B_ERR_NO_DEVICE, "libnccl.so.2 could not be loaded");
  sp_destroy(c);
  if (world == 1) return 0;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  F2B_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
  c->sp.comm = comm;
  c->sp.world = world;
  c->sp.rank = rank;
  c->sp.mode = c->option("sp_mode", 0);
  ++c->opt_gen;
  return 0;
}

}  // extern "C"
