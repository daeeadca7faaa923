// Block-range sharding of one long signal over the GPUs of a box: one process per GPU, each owning a
// contiguous range of blocks of the same ordered stream the reference processes hop by hop (Python/apvast.py:153-165).
//
//   apv_range_run            S1-S3 only over the halo blocks (the streaming state is a finite-memory function of the
//                            inputs, apvast.py:115-151), then the owned blocks through the pipelined multi-block path;
//                            rendered outputs and filters of the owned blocks stay in HBM.
//   apv_range_exchange_halo  the one piece of data that crosses a range boundary: the overlap-add tail G[:, H:, :]
//                            (apvast.py:455-465) left by the last block of rank g is sent device-to-device to rank
//                            g + 1 (ncclSend / ncclRecv over NVLink) and added to its first Nb/H - 1 output blocks.
//   apv_range_gather         outputs + filters of every range land in the HBM of the root rank (ncclSend / ncclRecv),
//                            then go to host memory there.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 already in the process, e.g. the one torch loaded, else the
// system one), so the library itself links nothing but the CUDA runtime.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "engine.cuh"

using namespace apv;

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi g_nccl;

int nccl_load() {
  if (g_nccl.lib) return OK;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy already mapped into the process
  if (!lib && getenv("APV_NCCL_LIB")) lib = dlopen(getenv("APV_NCCL_LIB"), RTLD_NOW);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW);
  if (!lib) return fail(ENCCL, "cannot load libnccl.so.2: %s", dlerror());
  NcclApi a;
  a.lib = lib;
#define SYM(field, name)                                                            \
  *(void**)(&a.field) = dlsym(lib, name);                                           \
  if (!a.field) return fail(ENCCL, "libnccl: symbol %s not found", name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
  SYM(GetVersion, "ncclGetVersion")
#undef SYM
  g_nccl = a;
  return OK;
}

#define APV_NCCL_TRY(expr)                                                                            \
  do {                                                                                                \
    ncclResult_t _r = (expr);                                                                         \
    if (_r != ncclSuccess) return fail(ENCCL, "%s -> %s", #expr, g_nccl.GetErrorString(_r));      \
  } while (0)

// send[z][v][l][t] = G[z][v][l][H + t], t < T = Nb - H      grid (2 V L), one CTA per overlap row
__global__ void tail_pack_kernel(const double* __restrict__ G, double* __restrict__ send, int Nb, int H) {
  const int T = Nb - H;
  const double* g = G + (size_t)blockIdx.x * Nb + H;
  double* s = send + (size_t)blockIdx.x * T;
  for (int t = threadIdx.x; t < T; t += blockDim.x) s[t] = g[t];
}

// out[b][z][v][h][l] += tail[z][v][l][b H + h]  for b H + h < T       grid (ceil(H L / 256), 2 V, nblk)
__global__ void tail_add_kernel(double* __restrict__ out, const double* __restrict__ tail, int V, int L, int H, int T) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= H * L) return;
  const int h = e / L, l = e - h * L;
  const int zv = blockIdx.y, b = blockIdx.z;
  const int t = b * H + h;
  if (t >= T) return;
  out[((size_t)b * 2 * V + zv) * H * L + e] += tail[((size_t)zv * L + l) * T + t];
}

size_t per_out(const Dims& D) { return 2 * (size_t)D.V * D.H * D.L; }
size_t per_w(const Dims& D) { return 2 * (size_t)D.V * D.n; }

}  // namespace

namespace apv {

void range_free(Handle& h) {
  void* ps[] = {h.rg_out, h.rg_w, h.rg_info, h.rg_tail_send, h.rg_tail_recv, h.gat_out, h.gat_w, h.rg_sig, h.gat_info};
  for (void* p : ps)
    if (p) cudaFree(p);
  h.rg_out = h.rg_w = h.rg_tail_send = h.rg_tail_recv = h.gat_out = h.gat_w = h.rg_sig = nullptr;
  h.rg_info = h.gat_info = nullptr;
  h.rg_cap = h.gat_cap = h.rg_owned = 0;
  h.rg_sig_cap = 0;
  if (h.comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)h.comm);
  h.comm = nullptr;
}

// Device buffers of a block range: `max_owned` blocks of outputs + filters; on the gathering rank room for
// `total_on_root` blocks (0 on the other ranks).
int range_alloc(Handle& h, int max_owned, int total_on_root) {
  const Dims& D = h.D;
  if (max_owned > h.rg_cap) {
    for (void* p : {(void*)h.rg_out, (void*)h.rg_w, (void*)h.rg_info})
      if (p) cudaFree(p);
    h.rg_out = h.rg_w = nullptr; h.rg_info = nullptr; h.rg_cap = 0;
    APV_CUDA_TRY(cudaMalloc((void**)&h.rg_out, (size_t)max_owned * per_out(D) * sizeof(double)));
    APV_CUDA_TRY(cudaMalloc((void**)&h.rg_w, (size_t)max_owned * per_w(D) * sizeof(double)));
    APV_CUDA_TRY(cudaMalloc((void**)&h.rg_info, (size_t)max_owned * 8 * sizeof(int)));
    APV_CUDA_TRY(cudaMemset(h.rg_info, 0, (size_t)max_owned * 8 * sizeof(int)));
    h.rg_cap = max_owned;
  }
  if (!h.rg_tail_send) {
    const size_t tail = 2 * (size_t)D.V * D.L * (size_t)(D.Nb - D.H);
    APV_CUDA_TRY(cudaMalloc((void**)&h.rg_tail_send, (tail ? tail : 1) * sizeof(double)));
    APV_CUDA_TRY(cudaMalloc((void**)&h.rg_tail_recv, (tail ? tail : 1) * sizeof(double)));
  }
  if (total_on_root > h.gat_cap) {
    for (void* p : {(void*)h.gat_out, (void*)h.gat_w, (void*)h.gat_info})
      if (p) cudaFree(p);
    h.gat_out = h.gat_w = nullptr; h.gat_info = nullptr; h.gat_cap = 0;
    APV_CUDA_TRY(cudaMalloc((void**)&h.gat_out, (size_t)total_on_root * per_out(D) * sizeof(double)));
    APV_CUDA_TRY(cudaMalloc((void**)&h.gat_w, (size_t)total_on_root * per_w(D) * sizeof(double)));
    APV_CUDA_TRY(cudaMalloc((void**)&h.gat_info, (size_t)total_on_root * 8 * sizeof(int)));
    h.gat_cap = total_on_root;
  }
  return OK;
}

}  // namespace apv

extern "C" {

int apv_comm_unique_id(void* id128) {
  if (!id128) return fail(EINVAL_, "null argument");
  APV_TRY(nccl_load());
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  APV_NCCL_TRY(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return OK;
}

int apv_comm_init(apv_handle* h, int rank, int nranks, const void* id128) {
  if (!h || nranks < 1 || rank < 0 || rank >= nranks) return fail(EINVAL_, "bad argument");
  DevGuard dg(h->device);
  if (h->comm) {
    g_nccl.CommDestroy((ncclComm_t)h->comm);
    h->comm = nullptr;
  }
  h->comm_rank = rank;
  h->comm_size = nranks;
  if (nranks == 1) return OK;
  if (!id128) return fail(EINVAL_, "null unique id");
  APV_TRY(nccl_load());
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  APV_NCCL_TRY(g_nccl.CommInitRank(&c, nranks, id, rank));
  h->comm = c;
  return OK;
}

int apv_comm_destroy(apv_handle* h) {
  if (!h) return fail(EINVAL_, "null argument");
  DevGuard dg(h->device);
  if (h->comm) {
    cudaStreamSynchronize(h->st);
    APV_NCCL_TRY(g_nccl.CommDestroy((ncclComm_t)h->comm));
  }
  h->comm = nullptr;
  h->comm_rank = 0;
  h->comm_size = 1;
  return OK;
}

int apv_nccl_version(int* version) {
  if (!version) return fail(EINVAL_, "null argument");
  APV_TRY(nccl_load());
  APV_NCCL_TRY(g_nccl.GetVersion(version));
  return OK;
}

int apv_range_reserve(apv_handle* h, int max_halo, int max_owned, int total_on_root) {
  if (!h || max_owned < 1 || max_halo < 0 || total_on_root < 0) return fail(EINVAL_, "bad argument");
  DevGuard dg(h->device);
  APV_TRY(range_alloc(*h, max_owned, total_on_root));
  APV_TRY(ensure_depth(*h, -1));
  const size_t need = 2 * (size_t)(max_halo + max_owned) * h->D.H;
  if (need > h->rg_sig_cap) {
    if (h->rg_sig) cudaFree(h->rg_sig);
    h->rg_sig = nullptr; h->rg_sig_cap = 0;
    APV_CUDA_TRY(cudaMalloc((void**)&h->rg_sig, need * sizeof(double)));
    h->rg_sig_cap = need;
  }
  return OK;
}

int apv_range_run(apv_handle* h, int n_halo, int n_owned, const double* in_A, const double* in_B, int inputs_on_device) {
  if (!h || n_halo < 0 || n_owned < 0 || !in_A || !in_B) return fail(EINVAL_, "bad argument");
  if (h->cfg.perceptual >= 2) return fail(EINVAL_, "block ranges need the on-device perceptual model");
  if (h->cfg.perceptual == 1 && !h->G2) return fail(EINVAL_, "perceptual model tables not set (apv_set_gain_table)");
  DevGuard dg(h->device);
  const Dims& D = h->D;
  const int nb = n_halo + n_owned;
  APV_TRY(apv_range_reserve(h, n_halo, n_owned > 0 ? n_owned : 1, h->gat_cap));
  const double *dA = in_A, *dB = in_B;
  if (!inputs_on_device) {
    const size_t cnt = (size_t)nb * D.H;
    APV_CUDA_TRY(cudaMemcpyAsync(h->rg_sig, in_A, cnt * sizeof(double), cudaMemcpyHostToDevice, h->st));
    APV_CUDA_TRY(cudaMemcpyAsync(h->rg_sig + cnt, in_B, cnt * sizeof(double), cudaMemcpyHostToDevice, h->st));
    dA = h->rg_sig;
    dB = h->rg_sig + cnt;
  }
  for (int t = 0; t < n_halo; ++t)
    APV_TRY(enqueue_front(*h, t, dA + (size_t)t * D.H, dB + (size_t)t * D.H, true));
  auto own = [&](const double* base, int b) { return base + (size_t)(n_halo + b) * D.H; };
  if (n_owned > 0) APV_TRY(enqueue_front(*h, 0, own(dA, 0), own(dB, 0), false));
  for (int b = 0; b < n_owned; ++b) {
    const BlockSink sink{h->rg_out + (size_t)b * per_out(D), nullptr, h->rg_w + (size_t)b * per_w(D), h->rg_info + (size_t)b * 8};
    APV_TRY(enqueue_back(*h, b, sink));
    if (b + 1 < n_owned) APV_TRY(enqueue_front(*h, b + 1, own(dA, b + 1), own(dB, b + 1), false));
  }
  h->rg_owned = n_owned;
  return leave_multiblock(*h);
}

int apv_range_exchange_halo(apv_handle* h) {
  if (!h) return fail(EINVAL_, "null argument");
  DevGuard dg(h->device);
  const Dims& D = h->D;
  const int T = D.Nb - D.H;
  if (h->comm_size <= 1 || T <= 0) return OK;
  if (!h->comm) return fail(ENCCL, "apv_range_exchange_halo: no communicator (apv_comm_init)");
  const int k1 = ceil_div(D.Nb, D.H) - 1;       // output blocks an overlap tail reaches into
  const bool sends = h->comm_rank + 1 < h->comm_size, recvs = h->comm_rank > 0;
  if (sends && h->rg_owned < k1)
    return fail(EINVAL_, "a block range must hold at least Nb/H - 1 = %d blocks for its overlap tail to be complete", k1);
  const size_t cnt = 2 * (size_t)D.V * D.L * T;
  if (sends) tail_pack_kernel<<<2 * D.V * D.L, 256, 0, h->st>>>(h->G, h->rg_tail_send, D.Nb, D.H);
  APV_NCCL_TRY(g_nccl.GroupStart());
  if (sends) APV_NCCL_TRY(g_nccl.Send(h->rg_tail_send, cnt, ncclDouble, h->comm_rank + 1, (ncclComm_t)h->comm, h->st));
  if (recvs) APV_NCCL_TRY(g_nccl.Recv(h->rg_tail_recv, cnt, ncclDouble, h->comm_rank - 1, (ncclComm_t)h->comm, h->st));
  APV_NCCL_TRY(g_nccl.GroupEnd());
  if (recvs && h->rg_owned > 0) {
    const int nblk = std::min(k1, h->rg_owned);
    tail_add_kernel<<<dim3(ceil_div(D.H * D.L, 256), 2 * D.V, nblk), 256, 0, h->st>>>(h->rg_out, h->rg_tail_recv, D.V, D.L, D.H, T);
  }
  APV_CUDA_TRY(cudaGetLastError());
  return OK;
}

// Debug / test hook: the packed overlap tail this rank would send, and adding a tail without a communicator
// (two ranks emulated on one GPU).
int apv_range_tail_get(apv_handle* h, double* tail_host) {
  if (!h || !tail_host) return fail(EINVAL_, "null argument");
  DevGuard dg(h->device);
  const Dims& D = h->D;
  const int T = D.Nb - D.H;
  const size_t cnt = 2 * (size_t)D.V * D.L * T;
  if (cnt == 0) return OK;
  APV_TRY(range_alloc(*h, h->rg_cap > 0 ? h->rg_cap : 1, h->gat_cap));
  tail_pack_kernel<<<2 * D.V * D.L, 256, 0, h->st>>>(h->G, h->rg_tail_send, D.Nb, D.H);
  APV_CUDA_TRY(cudaMemcpyAsync(tail_host, h->rg_tail_send, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  APV_CUDA_TRY(cudaStreamSynchronize(h->st));
  return OK;
}

int apv_range_tail_add(apv_handle* h, const double* tail_host) {
  if (!h || !tail_host) return fail(EINVAL_, "null argument");
  DevGuard dg(h->device);
  const Dims& D = h->D;
  const int T = D.Nb - D.H;
  const size_t cnt = 2 * (size_t)D.V * D.L * T;
  if (cnt == 0 || h->rg_owned == 0) return OK;
  APV_CUDA_TRY(cudaMemcpyAsync(h->rg_tail_recv, tail_host, cnt * sizeof(double), cudaMemcpyHostToDevice, h->st));
  const int nblk = std::min(ceil_div(D.Nb, D.H) - 1, h->rg_owned);
  tail_add_kernel<<<dim3(ceil_div(D.H * D.L, 256), 2 * D.V, nblk), 256, 0, h->st>>>(h->rg_out, h->rg_tail_recv, D.V, D.L, D.H, T);
  APV_CUDA_TRY(cudaGetLastError());
  APV_CUDA_TRY(cudaStreamSynchronize(h->st));
  return OK;
}

int apv_range_gather(apv_handle* h, int root, const int* counts, double* out_host, double* w_host) {
  if (!h || !counts) return fail(EINVAL_, "null argument");
  DevGuard dg(h->device);
  const Dims& D = h->D;
  const int size = h->comm_size, rank = h->comm_rank;
  if (root < 0 || root >= size) return fail(EINVAL_, "bad root");
  if (counts[rank] != h->rg_owned) return fail(EINVAL_, "counts[rank] = %d but the last range held %d blocks", counts[rank], h->rg_owned);
  const size_t po = per_out(D), pw = per_w(D);
  int total = 0;
  std::vector<int> off(size, 0);
  for (int r = 0; r < size; ++r) { off[r] = total; total += counts[r]; }
  if (size > 1 && !h->comm) return fail(ENCCL, "apv_range_gather: no communicator (apv_comm_init)");
  if (rank == root) {
    APV_TRY(range_alloc(*h, h->rg_cap > 0 ? h->rg_cap : 1, total));
    if (h->rg_owned > 0) {
      APV_CUDA_TRY(cudaMemcpyAsync(h->gat_out + (size_t)off[rank] * po, h->rg_out, (size_t)h->rg_owned * po * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
      APV_CUDA_TRY(cudaMemcpyAsync(h->gat_w + (size_t)off[rank] * pw, h->rg_w, (size_t)h->rg_owned * pw * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
      APV_CUDA_TRY(cudaMemcpyAsync(h->gat_info + (size_t)off[rank] * 8, h->rg_info, (size_t)h->rg_owned * 8 * sizeof(int), cudaMemcpyDeviceToDevice, h->st));
    }
  }
  if (size > 1) {
    ncclComm_t c = (ncclComm_t)h->comm;
    APV_NCCL_TRY(g_nccl.GroupStart());
    if (rank == root) {
      for (int r = 0; r < size; ++r) {
        if (r == root || counts[r] == 0) continue;
        APV_NCCL_TRY(g_nccl.Recv(h->gat_out + (size_t)off[r] * po, (size_t)counts[r] * po, ncclDouble, r, c, h->st));
        APV_NCCL_TRY(g_nccl.Recv(h->gat_w + (size_t)off[r] * pw, (size_t)counts[r] * pw, ncclDouble, r, c, h->st));
        APV_NCCL_TRY(g_nccl.Recv(h->gat_info + (size_t)off[r] * 8, (size_t)counts[r] * 8, ncclInt32, r, c, h->st));
      }
    } else if (h->rg_owned > 0) {
      APV_NCCL_TRY(g_nccl.Send(h->rg_out, (size_t)h->rg_owned * po, ncclDouble, root, c, h->st));
      APV_NCCL_TRY(g_nccl.Send(h->rg_w, (size_t)h->rg_owned * pw, ncclDouble, root, c, h->st));
      APV_NCCL_TRY(g_nccl.Send(h->rg_info, (size_t)h->rg_owned * 8, ncclInt32, root, c, h->st));
    }
    APV_NCCL_TRY(g_nccl.GroupEnd());
  }
  int rc = OK;
  if (rank == root) {
    std::vector<int> info((size_t)total * 8, 0);
    if (out_host && total > 0)
      APV_CUDA_TRY(cudaMemcpyAsync(out_host, h->gat_out, (size_t)total * po * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    if (w_host && total > 0)
      APV_CUDA_TRY(cudaMemcpyAsync(w_host, h->gat_w, (size_t)total * pw * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    if (total > 0)
      APV_CUDA_TRY(cudaMemcpyAsync(info.data(), h->gat_info, info.size() * sizeof(int), cudaMemcpyDeviceToHost, h->st));
    APV_CUDA_TRY(cudaStreamSynchronize(h->st));
    for (int b = 0; b < total && rc == OK; ++b) rc = status_from_info(*h, info.data() + (size_t)b * 8, b);
  } else {
    APV_CUDA_TRY(cudaStreamSynchronize(h->st));
  }
  APV_CUDA_TRY(cudaStreamSynchronize(h->st_front));
  return rc;
}

/* device pointers of the gathered results on the root (valid until the next apv_range_reserve / apv_destroy) */
int apv_range_device_ptrs(apv_handle* h, void** gathered_out, void** gathered_w, void** own_out, void** own_w) {
  if (!h) return fail(EINVAL_, "null argument");
  if (gathered_out) *gathered_out = h->gat_out;
  if (gathered_w) *gathered_w = h->gat_w;
  if (own_out) *own_out = h->rg_out;
  if (own_w) *own_w = h->rg_w;
  return OK;
}

void* apv_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    fail(ENOMEM_, "cudaMallocHost(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}

void apv_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"
