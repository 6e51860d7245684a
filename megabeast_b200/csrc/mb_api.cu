// megabeast_b200 host side of the C ABI (include/megabeast_b200.h): one-time ingest of the
// reference's data dictionaries into device-resident class-coded form, and the per-step launch
// sequence: K2 alone (FACTOR mode), K0 -> K1 -> K2 (TABLE modes) or K1e -> K1b -> K2
// (ELEMENTWISE); the reduction over stars (K3) is K2's tail.  No torch, no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "mb_kernels.cuh"

using namespace mb;

// host-side form of an entry while a star is being sorted / merged
struct Entry {
  double a;
  uint32_t idx;   // dust class (FACTOR), class combination (TABLE_UNIQUE) or grid row (TABLE_GRID)
  uint32_t meta;  // age class
};

static thread_local std::string g_err;
static int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
#define CK(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

static int nvars_of(const mb_model_desc& m, int p) {
  switch (m.kind[p]) {
    case MB_LOGNORMAL: return 2;
    case MB_TWO_LOGNORMAL: return 5;
    case MB_BINS_HISTO: case MB_BINS_INTERP: return m.nnodes[p];
    case MB_EXPONENTIAL: return 1;
    default: return 0;  // MB_FIXED, MB_FLAT, MB_TABULATED
  }
}

struct mb_handle {
  int device = 0, mode = 0, poisson = 0, ndim = 0, sm_count = 0, max_smem = 0;
  mb_model_desc model{};
  ModelDev mdev{};
  std::vector<double> clsx[MB_NPARAM];
  int nA = 1, nB = 1, nbins = 0, nseg[2] = {0, 0}, ncls_total = 0;
  // FACTOR-mode table split: outer table = age classes x the dust parameters kept outside (nO = nA * nOd
  // rows), inner table = the other dust parameters (nI rows).  Classic split: nOd = 1, nI = nB.
  int nO = 1, nOd = 1, nI = 1, inner_mask = 0, nscratch = 0, age_slot = 0;
  int ob = 0;                                // bits of outer class in a 16-bit code word (0: 32-bit codes)
  uint32_t oradix[MB_NPARAM] = {0, 0, 0, 0}, iradix[MB_NPARAM] = {0, 0, 0, 0};
  int64_t G = 0, S_local = 0, S_total = 0, n_entries = 0, n_entries_raw = 0, U = 0, n_runs = 0;
  int64_t device_bytes = 0;
  // parity hook (host copies)
  std::vector<int64_t> kept_rows, kept_offs;
  // device, static
  double* d_clsx = nullptr;
  double* d_ea = nullptr;       // entry weights
  uint32_t* d_ec = nullptr;     // entry codes (32-bit), or ...
  uint16_t* d_ec16 = nullptr;   // ... 16-bit codes: FACTOR mode, fp64 tables, <= 16 age classes (see Lanes)
  int c16 = 0;
  int64_t* d_seg[2] = {nullptr, nullptr};  // segment tables: [0] 768-thread kernels, [1] 512-thread kernel
  uint32_t* d_rowcode = nullptr;
  uint16_t* d_rowsrc = nullptr;   // [nO + nI][4] look-up rows multiplied into a table row (K0)
  uint16_t* d_rowslot = nullptr;  // the same as slots of K2's own table build (0xfffe: evaluated in place)
  uint16_t* d_lutslot = nullptr;  // [ncls_total] where a look-up row lives during K2's table build
  double* d_Pb = nullptr;         // [nbins][2][128] per-bin N-stars results of the running launch
  // ELEMENTWISE mode: the physics columns themselves, bin-sorted row list for the N-stars sums
  double* d_col[MB_NPARAM] = {nullptr, nullptr, nullptr, nullptr};
  double* d_gw = nullptr;
  double* d_compl = nullptr;
  int32_t* d_binrows = nullptr;      // grid rows in (age, Z)-bin order
  int64_t* d_tile_start = nullptr;   // K1e tiles over that list, never straddling a bin
  int64_t* d_bin_tile = nullptr;     // [nbins + 1] first tile of every bin
  double* d_partial = nullptr;       // [ntiles][128][2] per-tile partial bin sums
  int64_t ntiles = 0;
  double* d_bin_age = nullptr;
  double* d_Fage = nullptr;     // [nbins][128] age factor per bin, written by K1e
  double* d_SC = nullptr;       // [nbins][128][2] bin sums, written by K1b
  double* d_H = nullptr;
  double* d_coef = nullptr;
  int32_t* d_bin_agecls = nullptr;
  // device, per evaluation (sized for cap_wpad)
  double* d_phi = nullptr;  int64_t cap_phi = 0;
  double* d_tab[MB_NPARAM] = {nullptr, nullptr, nullptr, nullptr};
  int64_t cap_tab[MB_NPARAM] = {0, 0, 0, 0};
  double* d_F = nullptr;    // [nO + nI + nA][128]: outer rows, inner rows, plain age rows
  unsigned int* d_counter = nullptr;
  unsigned long long* d_gacc = nullptr;
  unsigned long long* d_timeline = nullptr;    // K2 phase stamps, allocated when profiling is on
  int timeline_blocks = 0;
  unsigned long long* d_xbuf = nullptr;        // this GPU's exchange buffer (cudaMalloc: IPC-exportable)
  unsigned long long* peer_xbuf[MB_MAX_PEERS] = {};
  int peer_world = 1, peer_rank = 0;
  unsigned int peer_seq = 0, rdv_seq = 0;
  double* d_T = nullptr;  int64_t cap_T = 0;
  int max_blocks = 0;
  double* d_out = nullptr;  int64_t cap_out = 0;
  double* h_pin = nullptr;  int64_t cap_pin = 0;  // mapped pinned host buffer: phi | result rows
  double* d_pin = nullptr;                         // its device address
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[MB_NKERNELS + 1] = {};
  int profiling = 0;
  float last_ms[MB_NKERNELS] = {};
  int last_wp = 0, last_launches = 0;
  int pending_W = 0;
  std::vector<int32_t> scratch_idx, scratch_st;
  std::vector<double> scratch_phi, scratch_rows, scratch_val;
  int table_bytes = 8;   // FACTOR-mode shared-memory table element: 8 = fp64, 4 = fp32
  int no_fuse = 0;       // keep K0 as a separate launch (MB_NO_FUSE=1: A/B measurements)
  // host path: K2's last block raises a flag in mapped host memory after the result rows; the host
  // polls it instead of going through cudaStreamSynchronize (MB_NO_FLAG=1: always synchronize)
  int use_flag = 1;
  unsigned long long flag_seq = 0;     // value the pending evaluation will publish (0: no flag armed)
  volatile unsigned long long* h_flag = nullptr;
  unsigned long long* d_flag = nullptr;
  std::vector<int64_t> star_end_pos;  // entry count after each non-empty star
  int spw = 1;
  double tail_frac = 0.0;
  unsigned long long peer_timeout_ns = 20000000000ull;  // MB_PEER_TIMEOUT_MS (default 20 s)
  int occ_cache[3] = {0, 0, 0};
  size_t occ_smem[3] = {0, 0, 0};
  // host arrays the caller asked us to watch for in-place changes (mb_watch_host)
  struct Watch { const void* ptr; int64_t nbytes; uint64_t fp; };
  std::vector<Watch> watches;
  // pixel batches
  int64_t P = 0;
  int64_t* d_pix_seg = nullptr;
  int32_t* d_pix_nstars = nullptr;
  double* d_Fpix = nullptr; int64_t cap_Fpix = 0;
  double* pending_out = nullptr;
};

template <typename T>
static int dev_alloc(mb_handle* h, T** p, int64_t n) {
  if (n <= 0) n = 1;
  CK(cudaMalloc((void**)p, sizeof(T) * (size_t)n));
  h->device_bytes += (int64_t)sizeof(T) * n;
  return 0;
}
template <typename T>
static int dev_upload(mb_handle* h, T** p, const std::vector<T>& v) {
  if (dev_alloc(h, p, (int64_t)v.size())) return 1;
  if (!v.empty()) CK(cudaMemcpy(*p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
  return 0;
}

// fingerprint of a host array: 16 eight-byte words spread evenly over it, plus its size (FNV-1a)
static uint64_t host_fingerprint(const void* ptr, int64_t nbytes) {
  uint64_t hsh = 1469598103934665603ull ^ (uint64_t)nbytes;
  const int64_t nw = nbytes / 8;
  if (nw <= 0) return hsh;
  const unsigned char* base = static_cast<const unsigned char*>(ptr);
  for (int i = 0; i < 16; ++i) {
    uint64_t w;
    memcpy(&w, base + 8 * ((nw - 1) * i / 15), 8);
    hsh = (hsh ^ w) * 1099511628211ull;
  }
  return hsh;
}
static int check_watches(const mb_handle* h) {
  for (size_t i = 0; i < h->watches.size(); ++i)
    if (host_fingerprint(h->watches[i].ptr, h->watches[i].nbytes) != h->watches[i].fp)
      return fail("MB_STALE: watched host array %zu changed in place since the device copy was made", i);
  return 0;
}

extern "C" int mb_watch_host(mb_handle* h, const void* ptr, int64_t nbytes) {
  if (!h) return fail("mb_watch_host: null handle");
  if (!ptr) { h->watches.clear(); return 0; }
  if (nbytes < 0) return fail("mb_watch_host: negative size");
  if (h->watches.size() >= 16) return fail("mb_watch_host: at most 16 arrays");
  h->watches.push_back({ptr, nbytes, host_fingerprint(ptr, nbytes)});
  return 0;
}

extern "C" int mb_abi_version(void) { return MB_ABI_VERSION; }
extern "C" const char* mb_last_error(void) { return g_err.c_str(); }
extern "C" int mb_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    fail("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return -1;
  }
  return n;
}

// class code of every grid row for one parameter
// Returns 0 on success, 1 on error, 2 when the column has more than kMaxClassesPerColumn distinct
// values (not a Cartesian axis: the caller switches to the ELEMENTWISE mode or reports it).
static const size_t kMaxClassesPerColumn = 4096;
static int classify(const mb_model_desc& m, int p, const double* col, int64_t G,
                    std::vector<double>& clsx, std::vector<int32_t>& code) {
  code.resize((size_t)G);
  if (m.kind[p] == MB_BINS_HISTO) {
    // f depends on x only through its interval: classes = nodes.  Outside the node range the
    // model itself refuses (beast raises ValueError("requested x outside of model x range")).
    int n = m.nnodes[p];
    clsx.assign(m.nodes[p], m.nodes[p] + n);
    for (int i = 1; i < n; ++i)
      if (!(clsx[i] > clsx[i - 1])) return fail("bins_histo nodes must be strictly increasing");
    for (int64_t g = 0; g < G; ++g) {
      double x = col[g];
      if (!(x >= clsx[0] && x <= clsx[n - 1]))
        return fail("requested x outside of model x range");
      int i = (int)(std::upper_bound(clsx.begin(), clsx.end(), x) - clsx.begin()) - 1;
      code[(size_t)g] = std::min(i, n - 1);
    }
    return 0;
  }
  // distinct values, kept sorted; BEAST columns hold tens of them
  std::vector<double> uniq;
  for (int64_t g = 0; g < G; ++g) {
    double x = col[g];
    if (x != x) return fail("NaN in a physics-parameter column (row %lld)", (long long)g);
    auto it = std::lower_bound(uniq.begin(), uniq.end(), x);
    if (it == uniq.end() || *it != x) {
      if (uniq.size() >= kMaxClassesPerColumn) {
        fail("parameter column %d has more than %zu distinct values: not class-codable, use the ELEMENTWISE mode",
             p, kMaxClassesPerColumn);
        return 2;
      }
      uniq.insert(it, x);
    }
  }
  for (int64_t g = 0; g < G; ++g)
    code[(size_t)g] = (int32_t)(std::lower_bound(uniq.begin(), uniq.end(), col[g]) - uniq.begin());
  clsx = uniq;
  return 0;
}

static size_t k1e_smem_bytes(int Wpad, int ndim) {
  return sizeof(double) * ((size_t)Wpad * std::max(1, ndim) + (size_t)MB_NPARAM * Wpad * kK1eWalkerVals +
                           2 * (size_t)MB_NPARAM * kK1eRows + 2 * (size_t)kK1eRows + 4 * (size_t)kK1eThreads) +
         sizeof(int) * ((size_t)MB_NPARAM * kK1eRows + kK1eRows) + 16;
}
// K2's dynamic shared memory: [front][factor tables (FACTOR mode)].  The front region holds the
// per-warp entry stage and is reused as scratch: the in-kernel table build (look-up rows, optional),
// the N-stars prologue (per-warp partial bin sums) and the block-reduction slabs -- which move over
// the tables (dead after the star loop) when the tables are larger than they are, so that they never
// size the front.
struct K2Smem {
  size_t front = 0, tables = 0, total = 0, slab_off = 0;
  size_t scratch_off = 0;   // table-build / N-stars scratch: 0 (shares the stage) or right after the stage
  size_t pois_off = 0;      // N-stars scratch: after the look-up rows when the tables are built in the kernel
  size_t hts_off = 0;       // one-warp-per-bin N-stars scheme: staged statistics
  int hts_slices = 0;
  int pois_warps = 0;       // block-wide N-stars scheme: warps whose partial sums fit
  bool separate = false;    // stage and scratch disjoint: the entry stage is primed before the table build
  bool warp_private = false;// N-stars term: one warp per bin (few dust classes)
  bool fuse_fits = false;   // the scratch of the in-kernel table build fits as well
};
// scratch of the in-kernel table build: look-up rows | ln x, 1/x per class | per-walker constants |
// row lists | slots | phi
static size_t k2_fuse_scratch(int nscratch, int ncls_total, int nrows, int Wpad, size_t phi_doubles) {
  return 8 * ((size_t)nscratch * Wpad + 3 * (size_t)ncls_total + 6 * (size_t)MB_NPARAM * Wpad + (size_t)nrows +
              (size_t)((ncls_total + 3) / 4) + phi_doubles);
}
constexpr int kPoisPrivateMaxClasses = 512;  // one warp per bin up to this many dust classes
// Layouts are tried in order of preference: tables built in the kernel before a separate K0 launch,
// a private scratch (stage primed early, no barrier around the N-stars bins) before a shared one.
// nD = dust classes per bin, bins_per_block = most bins a block can get (0: no one-warp-per-bin scheme).
static K2Smem k2_smem_plan(int kmode, int nO, int nI, int nbins, int Wpad, int elem_bytes, int max_smem,
                           size_t fuse_scratch = 0, int nscratch = 0, int nD = 0, int bins_per_block = 0,
                           bool have_sc = false) {
  const int NW = Wpad / 32, threads = k2_threads(NW), warps = threads / 32;
  const size_t stage = (size_t)warps * k2_stage_bytes(NW);
  const size_t slabs = (size_t)warps * kAccRows * Wpad * 8;  // block reduction: one slab per warp
  const size_t pois_row = kPoisBins * 2 * (size_t)Wpad * 8;  // one warp's partial (S, C) of the bins in flight, all walkers
  const size_t tables = kmode == kModeFactor ? (size_t)(nO + nI) * Wpad * elem_bytes : 0;
  const bool slabs_over_tables = tables >= slabs;
  static const bool no_separate = getenv("MB_NO_PRIME") && atoi(getenv("MB_NO_PRIME"));        // A/B switches
  static const bool no_private = getenv("MB_NO_WARP_BINS") && atoi(getenv("MB_NO_WARP_BINS"));  // (measurements)
  auto layout = [&](bool sep, bool fused, bool priv, K2Smem& p) -> bool {
    p = K2Smem();
    p.tables = tables;
    p.separate = sep;
    p.fuse_fits = fused;
    p.warp_private = priv;
    p.scratch_off = sep ? ((stage + 15) & ~(size_t)15) : 0;
    p.pois_off = p.scratch_off + (fused ? (size_t)nscratch * Wpad * 8 : 0);
    size_t front = std::max(stage, p.scratch_off + (fused ? fuse_scratch : 0));
    if (!slabs_over_tables) front = std::max(front, slabs);
    if (nbins > 0 && !priv) {
      if (tables + p.pois_off + pois_row > (size_t)max_smem) return false;
      const size_t avail = (size_t)max_smem - tables - p.pois_off;
      p.pois_warps = (int)std::min<size_t>((size_t)warps, avail / pois_row);
      if (sep && p.pois_warps < warps) return false;  // a private scratch is only worth it at full width
      front = std::max(front, p.pois_off + (size_t)p.pois_warps * pois_row);
    }
    front = (front + 15) & ~(size_t)15;
    if (nbins > 0 && !have_sc && nD > 0 && bins_per_block > 0) {
      // the statistics of the block's bins staged in shared memory (cp.async at kernel start), after
      // everything else: the one-warp-per-bin scheme needs all of them, the block-wide one takes what fits
      const size_t slice = (size_t)nD * 16;
      const size_t room = tables + front < (size_t)max_smem ? (size_t)max_smem - tables - front : 0;
      p.hts_off = front;
      p.hts_slices = (int)std::min<size_t>((size_t)std::min(bins_per_block, 64), room / slice);
      if (priv && p.hts_slices < std::min(bins_per_block, warps)) return false;
      front += (size_t)p.hts_slices * slice;
    }
    p.front = (front + 15) & ~(size_t)15;
    p.slab_off = slabs_over_tables ? p.front : 0;
    p.total = p.front + p.tables;
    return p.total <= (size_t)max_smem;
  };
  const bool private_ok = !no_private && !no_separate && nbins > 0 && bins_per_block > 0 &&
                          (have_sc || (nD > 0 && nD <= kPoisPrivateMaxClasses));
  K2Smem p;
  for (int fused = fuse_scratch > 0 ? 1 : 0; fused >= 0; --fused) {
    if (private_ok && layout(true, fused != 0, true, p)) return p;
    for (int sep = no_separate ? 0 : 1; sep >= 0; --sep)
      if (layout(sep != 0, fused != 0, false, p)) return p;
  }
  p.total = (size_t)max_smem + 1;  // nothing fits
  return p;
}

extern "C" int mb_peer_disconnect(mb_handle* h) {
  if (!h) return fail("null handle");
  CK(cudaSetDevice(h->device));
  for (int p = 0; p < h->peer_world; ++p)
    if (p != h->peer_rank && h->peer_xbuf[p]) cudaIpcCloseMemHandle(h->peer_xbuf[p]);
  for (auto& x : h->peer_xbuf) x = nullptr;
  h->peer_world = 1;
  h->peer_rank = 0;
  return 0;
}

extern "C" int mb_peer_export(mb_handle* h, void* token) {
  if (!h || !token) return fail("mb_peer_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) <= MB_PEER_HANDLE_BYTES, "token too small");
  CK(cudaSetDevice(h->device));
  cudaIpcMemHandle_t ipc;
  CK(cudaIpcGetMemHandle(&ipc, h->d_xbuf));
  memset(token, 0, MB_PEER_HANDLE_BYTES);
  memcpy(token, &ipc, sizeof ipc);
  return 0;
}

extern "C" int mb_peer_connect(mb_handle* h, int32_t rank, int32_t world, const void* tokens) {
  if (!h || !tokens) return fail("mb_peer_connect: null argument");
  if (world < 1 || world > MB_MAX_PEERS || rank < 0 || rank >= world)
    return fail("mb_peer_connect: rank %d of %d (at most %d ranks)", rank, world, MB_MAX_PEERS);
  if (h->peer_world > 1 && mb_peer_disconnect(h)) return 1;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemset(h->d_xbuf, 0, kXBufBytes));
  for (int p = 0; p < world; ++p) {
    if (p == rank) { h->peer_xbuf[p] = h->d_xbuf; continue; }
    cudaIpcMemHandle_t ipc;
    memcpy(&ipc, (const char*)tokens + (size_t)p * MB_PEER_HANDLE_BYTES, sizeof ipc);
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, ipc, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (int q = 0; q < p; ++q)
        if (q != rank && h->peer_xbuf[q]) cudaIpcCloseMemHandle(h->peer_xbuf[q]);
      for (auto& x : h->peer_xbuf) x = nullptr;
      return fail("cudaIpcOpenMemHandle for rank %d failed: %s", p, cudaGetErrorString(e));
    }
    h->peer_xbuf[p] = (unsigned long long*)ptr;
  }
  h->peer_world = world;
  h->peer_rank = rank;
  h->peer_seq = 0;
  h->rdv_seq = 0;
  return 0;
}

extern "C" int mb_peer_rendezvous(mb_handle* h, void* stream) {
  if (!h) return fail("null handle");
  if (h->peer_world <= 1) return 0;
  CK(cudaSetDevice(h->device));
  PeerArgs pa{};
  pa.world = h->peer_world; pa.rank = h->peer_rank; pa.seq = ++h->rdv_seq; pa.timeout_ns = h->peer_timeout_ns;
  for (int p = 0; p < h->peer_world; ++p) pa.xbuf[p] = h->peer_xbuf[p];
  mb_rendezvous<<<1, 32, 0, (cudaStream_t)stream>>>(pa);
  CK(cudaGetLastError());
  return 0;
}

extern "C" void mb_destroy(mb_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->peer_world > 1) mb_peer_disconnect(h);
  if (h->d_xbuf) cudaFree(h->d_xbuf);
  void* ptrs[] = {h->d_clsx, h->d_ea, h->d_ec, h->d_ec16, h->d_seg[0], h->d_seg[1], h->d_rowcode, h->d_H, h->d_coef,
                  h->d_bin_agecls, h->d_phi, h->d_tab[0], h->d_tab[1], h->d_tab[2], h->d_tab[3],
                  h->d_F, h->d_counter, h->d_gacc, h->d_T, h->d_out, h->d_pix_seg, h->d_pix_nstars, h->d_Fpix,
                  h->d_timeline, h->d_col[0], h->d_col[1], h->d_col[2], h->d_col[3], h->d_gw, h->d_compl,
                  h->d_rowsrc, h->d_rowslot, h->d_lutslot, h->d_Pb, h->d_binrows, h->d_tile_start, h->d_bin_tile, h->d_partial, h->d_bin_age, h->d_Fage, h->d_SC};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  for (auto& e : h->ev)
    if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

// Star-aligned segments of the entry stream, one per resident warp of block shape t (0: the
// 768-thread kernels, 1: the 512-thread kernel), equal entry counts: boundary i is the first
// star end at or after its goal.  (Every block does the same prologue work now -- the N-stars bins
// are spread over all blocks -- so all warps get equal shares.)
static void build_segments(const mb_handle* h, int t, std::vector<int64_t>& seg) {
  const std::vector<int64_t>& ends = h->star_end_pos;
  const int threads = k2_threads(t == 0 ? 2 : 4), kw = threads / 32;
  const int64_t resident_warps = (int64_t)h->sm_count * kw;
  const int64_t n = h->n_entries;
  int64_t want = std::max<int64_t>(1, std::min<int64_t>(resident_warps * h->spw, n / 64));
  const int64_t n_main = want >= resident_warps ? (int64_t)((1.0 - h->tail_frac) * (double)n) : n;
  const int64_t want_tail = n_main < n ? std::max<int64_t>(1, (n - n_main) / 96) : 0;
  seg.assign(1, 0);
  size_t j = 0;
  for (int64_t i = 1; i <= want + want_tail && j < ends.size(); ++i) {
    int64_t goal;
    if (i > want) goal = n_main + ((n - n_main) * (i - want) + want_tail - 1) / want_tail;
    else goal = (n_main * i + want - 1) / want;
    while (j < ends.size() && ends[j] < goal) ++j;
    if (j < ends.size() && ends[j] > seg.back()) seg.push_back(ends[j]);
  }
  if (n > 0 && seg.back() != n) seg.push_back(n);
}

// Entry build on the device (mb_build_entries + mb_compact_entries): uploads this handle's columns of
// the sparse arrays and the three static grid columns, leaves the packed entry stream in d_ea / d_ec
// (or d_ec16) and the per-star bookkeeping the segment and pixel tables need on the host.
static int build_entries_device(mb_handle* h, const mb_grid_desc* grid, const mb_stars_desc* stars, int64_t s0,
                                int64_t s1, const uint32_t* rowcode, int mode, bool merge, int64_t n_pixels,
                                const int64_t* pix_off, std::vector<int64_t>& star_end_pos,
                                std::vector<int64_t>& pix_ent, std::vector<int64_t>& pix_se) {
  const int64_t K = stars->K, S = stars->S, Sl = s1 - s0, G = grid->G;
  h->n_entries = h->n_entries_raw = h->n_runs = 0;
  struct Temps {
    std::vector<void*> p;
    ~Temps() { for (void* q : p) cudaFree(q); }
    int get(void** out, size_t bytes) {
      *out = nullptr;
      cudaError_t e = cudaMalloc(out, std::max<size_t>(bytes, 16));
      if (e != cudaSuccess) return fail("cudaMalloc of %zu bytes for the entry build failed: %s", bytes, cudaGetErrorString(e));
      p.push_back(*out);
      return 0;
    }
  } tmp;
  std::vector<int32_t> cnt((size_t)Sl, 0), nraw((size_t)Sl, 0), nrun((size_t)Sl, 0);
  int32_t* d_cnt = nullptr; double* d_ta = nullptr; uint32_t* d_tc = nullptr;
  if (Sl > 0 && K > 0) {
    int64_t* d_idx; double* d_val; uint32_t* d_rc = nullptr; double *d_pw, *d_gw, *d_cm;
    int32_t *d_nraw, *d_nrun, *d_err;
    if (tmp.get((void**)&d_idx, sizeof(int64_t) * (size_t)(K * Sl)) || tmp.get((void**)&d_val, sizeof(double) * (size_t)(K * Sl)) ||
        tmp.get((void**)&d_pw, sizeof(double) * (size_t)G) || tmp.get((void**)&d_gw, sizeof(double) * (size_t)G) ||
        tmp.get((void**)&d_cm, sizeof(double) * (size_t)G) || tmp.get((void**)&d_ta, sizeof(double) * (size_t)(K * Sl)) ||
        tmp.get((void**)&d_tc, sizeof(uint32_t) * (size_t)(K * Sl)) || tmp.get((void**)&d_cnt, sizeof(int32_t) * (size_t)Sl) ||
        tmp.get((void**)&d_nraw, sizeof(int32_t) * (size_t)Sl) || tmp.get((void**)&d_nrun, sizeof(int32_t) * (size_t)Sl) ||
        tmp.get((void**)&d_err, sizeof(int32_t)))
      return 1;
    if (rowcode) {
      if (tmp.get((void**)&d_rc, sizeof(uint32_t) * (size_t)G)) return 1;
      CK(cudaMemcpy(d_rc, rowcode, sizeof(uint32_t) * (size_t)G, cudaMemcpyHostToDevice));
    }
    // this handle's columns [s0, s1) of the (K, S) arrays
    CK(cudaMemcpy2D(d_idx, sizeof(int64_t) * (size_t)Sl, stars->indxs + s0, sizeof(int64_t) * (size_t)S,
                    sizeof(int64_t) * (size_t)Sl, (size_t)K, cudaMemcpyHostToDevice));
    CK(cudaMemcpy2D(d_val, sizeof(double) * (size_t)Sl, stars->vals + s0, sizeof(double) * (size_t)S,
                    sizeof(double) * (size_t)Sl, (size_t)K, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_pw, grid->prior_weight, sizeof(double) * (size_t)G, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_gw, grid->grid_weight, sizeof(double) * (size_t)G, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_cm, grid->completeness, sizeof(double) * (size_t)G, cudaMemcpyHostToDevice));
    const int32_t no_err = 0x7fffffff;
    CK(cudaMemcpy(d_err, &no_err, sizeof no_err, cudaMemcpyHostToDevice));
    BuildArgs ba{};
    ba.indxs = d_idx; ba.vals = d_val; ba.K = K; ba.S = Sl; ba.G = G; ba.rowcode = d_rc;
    ba.pw = d_pw; ba.gw = d_gw; ba.cm = d_cm; ba.mode = mode; ba.nB = h->nB; ba.merge = merge ? 1 : 0;
    ba.tmp_a = d_ta; ba.tmp_c = d_tc; ba.cnt = d_cnt; ba.nraw = d_nraw; ba.nrun = d_nrun; ba.err_star = d_err;
    const int blocks = (int)std::min<int64_t>((Sl + kBuildWarps - 1) / kBuildWarps, (int64_t)h->sm_count * 32);
    if (K <= 32) mb_build_entries<32><<<blocks, 32 * kBuildWarps, 0, h->stream>>>(ba);
    else if (K <= 64) mb_build_entries<64><<<blocks, 32 * kBuildWarps, 0, h->stream>>>(ba);
    else mb_build_entries<128><<<blocks, 32 * kBuildWarps, 0, h->stream>>>(ba);
    CK(cudaGetLastError());
    int32_t err = no_err;
    CK(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(int32_t) * (size_t)Sl, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(nraw.data(), d_nraw, sizeof(int32_t) * (size_t)Sl, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(nrun.data(), d_nrun, sizeof(int32_t) * (size_t)Sl, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(&err, d_err, sizeof err, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (err != no_err) {  // report as numpy would: the first offending point of the first offending star
      const int64_t s = s0 + err;
      for (int64_t kk = 0; kk < K; ++kk) {
        if (!std::isfinite(stars->vals[kk * S + s])) continue;
        const int64_t r = stars->indxs[kk * S + s];
        if (r < -G || r >= G)
          return fail("index %lld is out of bounds for axis 0 with size %lld (star %lld)", (long long)r, (long long)G, (long long)s);
      }
      return fail("index out of bounds for axis 0 with size %lld (star %lld)", (long long)G, (long long)s);
    }
  }
  // prefix sums: entry offsets per star, ends of the non-empty stars, pixel bookkeeping
  std::vector<int64_t> offs((size_t)Sl + 1, 0), nonempty_before((size_t)Sl + 1, 0);
  for (int64_t s = 0; s < Sl; ++s) {
    offs[(size_t)s + 1] = offs[(size_t)s] + cnt[(size_t)s];
    nonempty_before[(size_t)s + 1] = nonempty_before[(size_t)s] + (cnt[(size_t)s] > 0 ? 1 : 0);
    if (cnt[(size_t)s] > 0) star_end_pos.push_back(offs[(size_t)s + 1]);
    h->n_entries_raw += nraw[(size_t)s];
    h->n_runs += nrun[(size_t)s];
  }
  h->n_entries = offs[(size_t)Sl];
  for (int64_t p = 0; p <= n_pixels && n_pixels > 0; ++p) {
    pix_ent.push_back(offs[(size_t)(pix_off[p] - s0)]);
    pix_se.push_back(nonempty_before[(size_t)(pix_off[p] - s0)]);
  }
  // (padded: the ring copies whole chunks, so it reads up to a chunk past the last entry)
  if (dev_alloc(h, &h->d_ea, h->n_entries + kEntryPad)) return 1;
  CK(cudaMemsetAsync(h->d_ea + h->n_entries, 0, sizeof(double) * kEntryPad, h->stream));
  if (h->c16) {
    if (dev_alloc(h, &h->d_ec16, h->n_entries + kEntryPad)) return 1;
    CK(cudaMemsetAsync(h->d_ec16 + h->n_entries, 0, sizeof(uint16_t) * kEntryPad, h->stream));
  } else {
    if (dev_alloc(h, &h->d_ec, h->n_entries + kEntryPad)) return 1;
    CK(cudaMemsetAsync(h->d_ec + h->n_entries, 0, sizeof(uint32_t) * kEntryPad, h->stream));
  }
  if (h->n_entries > 0) {
    int64_t* d_offs;
    if (tmp.get((void**)&d_offs, sizeof(int64_t) * offs.size())) return 1;
    CK(cudaMemcpyAsync(d_offs, offs.data(), sizeof(int64_t) * offs.size(), cudaMemcpyHostToDevice, h->stream));
    const int blocks = (int)std::min<int64_t>((Sl * 32 + 255) / 256, (int64_t)h->sm_count * 16);
    mb_compact_entries<<<blocks, 256, 0, h->stream>>>(d_ta, d_tc, d_cnt, d_offs, Sl, K, h->d_ea, h->d_ec, h->d_ec16, h->ob);
    CK(cudaGetLastError());
  }
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// returned by create_impl when the class-coded plan turned out impossible late and mode AUTO may still take
// the elementwise evaluation: mb_create then tries again with that mode
constexpr int kRetryElementwise = 3;
static int create_impl(const mb_grid_desc* grid, const mb_stars_desc* stars,
                       const mb_model_desc* model, const mb_massmult_desc* mm,
                       const mb_options* opts, int64_t n_pixels, const int64_t* pix_off, mb_handle** out) {
  if (!grid || !stars || !model || !out) return fail("mb_create: null argument");
  *out = nullptr;
  mb_options o{};
  if (opts) o = *opts;
  int ndev = 0;
  cudaError_t e0 = cudaGetDeviceCount(&ndev);
  if (e0 != cudaSuccess || ndev == 0)
    return fail("no CUDA device available (%s); megabeast_b200 has no CPU fallback",
                e0 == cudaSuccess ? "device count 0" : cudaGetErrorString(e0));
  if (o.device < 0 || o.device >= ndev) return fail("device %d out of range (0..%d)", o.device, ndev - 1);
  CK(cudaSetDevice(o.device));

  mb_handle* h = new mb_handle();
  struct Guard { mb_handle* h; bool ok = false; ~Guard() { if (!ok) mb_destroy(h); } } guard{h};
  h->device = o.device;
  h->model = *model;
  if (o.table_precision != 0 && o.table_precision != 1) return fail("table_precision must be 0 (fp64) or 1 (fp32)");
  h->table_bytes = o.table_precision == 1 ? 4 : 8;
  if (const char* nf = getenv("MB_NO_FUSE")) h->no_fuse = atoi(nf);
  if (const char* nf = getenv("MB_NO_FLAG")) h->use_flag = !atoi(nf);
  CK(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, o.device));
  CK(cudaDeviceGetAttribute(&h->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, o.device));
  h->max_smem -= 1024;  // room for the kernels' few bytes of static shared memory
  const int64_t G = grid->G;
  if (G <= 0 || G > 0x7fffffffLL) return fail("grid size %lld unsupported", (long long)G);
  h->G = G;
  if (!grid->prior_weight || !grid->grid_weight || !grid->completeness)
    return fail("prior_weight, grid_weight and completeness columns are required");

  // ---- phi layout ---------------------------------------------------------------------------------
  ModelDev& md = h->mdev;
  int k = 0;
  for (int p = 0; p < MB_NPARAM; ++p) {
    if (model->kind[p] < MB_FIXED || model->kind[p] > MB_TABULATED)
      return fail("unknown prior model kind %d for parameter %d", model->kind[p], p);
    if ((model->kind[p] == MB_BINS_HISTO || model->kind[p] == MB_BINS_INTERP) &&
        (model->nnodes[p] < 1 || model->nnodes[p] > MB_MAX_NODES))
      return fail("parameter %d: %d nodes (1..%d supported)", p, model->nnodes[p], MB_MAX_NODES);
    md.kind[p] = model->kind[p];
    md.nnodes[p] = model->nnodes[p];
    md.voff[p] = k;
    memcpy(md.nodes[p], model->nodes[p], sizeof(double) * MB_MAX_NODES);
    k += nvars_of(*model, p);
  }
  h->ndim = k;
  h->poisson = (mm != nullptr);
  if (h->poisson && model->kind[MB_P_LOGA] == MB_FIXED)
    return fail("mass multipliers given but the logA prior is fixed (no N-stars term, :71-74)");

  // ---- class codes per grid row ------------------------------------------------------------------
  // ELEMENTWISE (requested, or AUTO on a grid whose columns are not class-codable): no class codes
  // at all, the prior models are evaluated on every grid row (mb_k1_elementwise)
  bool elementwise = o.mode == MB_MODE_ELEMENTWISE;
  std::vector<int32_t> code[MB_NPARAM];
  int clsoff = 0;
  for (int p = 0; p < MB_NPARAM; ++p) {
    md.ncls[p] = 1;
    md.clsoff[p] = clsoff;
    md.radix[p] = 0;
    if (model->kind[p] == MB_FIXED) continue;
    if (!grid->col[p]) return fail("column of free parameter %d is NULL", p);
    if (elementwise) continue;
    const int rc = classify(*model, p, grid->col[p], G, h->clsx[p], code[p]);
    if (rc == 2 && o.mode == MB_MODE_AUTO && n_pixels == 0) { elementwise = true; continue; }
    if (rc) return 1;
    md.ncls[p] = (int)h->clsx[p].size();
    clsoff += md.ncls[p];
  }
  int64_t nB = 1;
  if (!elementwise) {
    for (int p = MB_P_AV; p < MB_NPARAM; ++p) {
      if (model->kind[p] == MB_FIXED) continue;
      md.radix[p] = (int)nB;
      nB *= md.ncls[p];
      if (nB > (int64_t)kDustMask + 1) {
        if (o.mode == MB_MODE_AUTO && n_pixels == 0) { elementwise = true; break; }
        return fail("more than 2^20 distinct dust-parameter combinations");
      }
    }
    if (clsoff > 4096 && !elementwise) {
      if (o.mode == MB_MODE_AUTO && n_pixels == 0) elementwise = true;
      else return fail("more than 4096 classes in total (%d): K0 look-up table too large", clsoff);
    }
  }
  if (elementwise) {
    if (n_pixels > 0) return fail("pixel batches need class-codable columns (FACTOR mode)");
    for (int p = 0; p < MB_NPARAM; ++p) {
      md.ncls[p] = 1; md.clsoff[p] = 0; md.radix[p] = 0; h->clsx[p].clear(); code[p].clear();
      if (model->kind[p] == MB_FIXED) continue;
      if (model->kind[p] == MB_TABULATED) {
        // a host-evaluated model on a column that cannot be class-coded: every grid row is its own class --
        // mb_class_values returns the column in row order, the table is [W][G] (the reference's own cost,
        // singlepop_dust_model.py:146-148).  Dust parameters only: the age factor of an (age, Z) bin is
        // needed apart from the rows (helpers.py:135)
        if (p == MB_P_LOGA) return fail("a tabulated age model needs a class-codable logA column (not available in ELEMENTWISE mode)");
        if (G > 0x7fffffffLL) return fail("grid too large for a per-row table");
        md.ncls[p] = (int32_t)G;
        h->clsx[p].assign(grid->col[p], grid->col[p] + G);
        continue;
      }
      const double* col = grid->col[p];
      const int n = model->nnodes[p];
      for (int64_t g = 0; g < G; ++g) {
        if (col[g] != col[g]) return fail("NaN in a physics-parameter column (row %lld)", (long long)g);
        if (model->kind[p] == MB_BINS_HISTO && !(col[g] >= model->nodes[p][0] && col[g] <= model->nodes[p][n - 1]))
          return fail("requested x outside of model x range");
      }
      if (model->kind[p] == MB_BINS_HISTO)
        for (int i = 1; i < n; ++i)
          if (!(model->nodes[p][i] > model->nodes[p][i - 1])) return fail("bins_histo nodes must be strictly increasing");
    }
    clsoff = 0;
    nB = 1;
  }
  h->ncls_total = clsoff;
  h->nB = (int)nB;
  h->nA = md.ncls[MB_P_LOGA];
  if (h->nA > (int)kAgeMask + 1) return fail("more than 1024 distinct age classes");
  // ---- mode, and the split of the factor tables ----------------------------------------------------
  // FACTOR keeps two factor tables in shared memory: outer (age classes x the dust parameters kept
  // outside, one row read per RUN of a star's entries) and inner (the other dust parameters, one row
  // read per ENTRY).  The classic split -- outer = age, inner = product over all dust parameters --
  // gives the fewest runs and is taken whenever its tables fit; production dust axes (100-200 A(V)
  // values x 9 R(V) ...) do not fit as one product table, so some dust parameters move to the outer
  // table: among the splits that fit, the one with the fewest outer rows (fewest runs per star).
  const int nbins_plan = mm ? mm->n_ages * mm->n_mets : 0;
  auto fits = [&](int64_t nO, int64_t nI, int Wpad) {
    return nO <= (int64_t)kAgeMask + 1 && nI <= (int64_t)kDustMask + 1 &&
           k2_smem_plan(kModeFactor, (int)nO, (int)nI, nbins_plan, Wpad, h->table_bytes, h->max_smem).total <=
               (size_t)h->max_smem;
  };
  std::vector<int> dust_free;
  for (int p = MB_P_AV; p < MB_NPARAM; ++p)
    if (model->kind[p] != MB_FIXED && !elementwise) dust_free.push_back(p);
  int full_mask = 0;
  for (int p : dust_free) full_mask |= 1 << (p - MB_P_AV);
  auto split_sizes = [&](int inner_mask, int64_t& nOd, int64_t& nI) {
    nOd = 1; nI = 1;
    for (int p : dust_free) (((inner_mask >> (p - MB_P_AV)) & 1) ? nI : nOd) *= md.ncls[p];
  };
  int inner_mask = full_mask;   // classic
  int mode = elementwise ? MB_MODE_ELEMENTWISE : o.mode;
  if (mode < MB_MODE_AUTO || mode > MB_MODE_ELEMENTWISE) return fail("unknown mode %d", mode);
  const int forced_mask = o.split_inner_mask & 7;
  if (forced_mask && (forced_mask & ~full_mask))
    return fail("split_inner_mask 0x%x names a dust parameter whose prior is fixed", forced_mask);
  if (forced_mask && n_pixels > 0) return fail("pixel batches use the classic table split");
  if ((mode == MB_MODE_AUTO || mode == MB_MODE_FACTOR) && !elementwise) {
    bool found = false;
    if (forced_mask) {
      int64_t nOd, nI;
      split_sizes(forced_mask, nOd, nI);
      found = fits(h->nA * nOd, nI, 32);
      inner_mask = forced_mask;
    } else {
      // walkers per pass the tables must fit for: 64 first (the usual batch), then 32
      for (int Wpad : {64, 32}) {
        if (fits(h->nA, h->nB, Wpad)) { inner_mask = full_mask; found = true; break; }
        if (n_pixels > 0) continue;
        int64_t best_nO = -1, best_rows = 0;
        for (int m = 1; m < full_mask; ++m) {
          if (m & ~full_mask) continue;
          int64_t nOd, nI;
          split_sizes(m, nOd, nI);
          const int64_t nO = h->nA * nOd;
          if (!fits(nO, nI, Wpad)) continue;
          if (best_nO < 0 || nO < best_nO || (nO == best_nO && nO + nI < best_rows)) {
            best_nO = nO; best_rows = nO + nI; inner_mask = m;
          }
        }
        if (best_nO >= 0) { found = true; break; }
      }
    }
    if (found) mode = MB_MODE_FACTOR;
    else if (mode == MB_MODE_FACTOR || forced_mask)
      return fail("factor tables (%d age x %d dust classes) do not fit shared memory in any split; use a TABLE mode",
                  h->nA, h->nB);
    else { mode = MB_MODE_TABLE_UNIQUE; inner_mask = full_mask; }
  }
  if (mode != MB_MODE_FACTOR) inner_mask = full_mask;
  if (mode == MB_MODE_TABLE_UNIQUE && (int64_t)h->nA * h->nB > (1LL << 27)) mode = MB_MODE_TABLE_GRID;
  h->mode = mode;
  h->U = (mode == MB_MODE_TABLE_GRID || elementwise) ? G : (int64_t)h->nA * h->nB;
  {
    int64_t nOd = 1, nI = 1;
    h->inner_mask = inner_mask;
    for (int p : dust_free) {
      if ((inner_mask >> (p - MB_P_AV)) & 1) { h->iradix[p] = (uint32_t)nI; nI *= md.ncls[p]; }
      else { h->oradix[p] = (uint32_t)nOd; nOd *= md.ncls[p]; }
    }
    h->nOd = (int)nOd; h->nI = (int)nI; h->nO = h->nA * (int)nOd;
  }
  // grid-row class code: outer class (age class * nOd + outer-dust digit) << 20 | inner class
  std::vector<uint32_t> rowcode(elementwise ? 0 : (size_t)G);
  for (int64_t g = 0; g < G && !elementwise; ++g) {
    uint32_t cA = model->kind[MB_P_LOGA] == MB_FIXED ? 0u : (uint32_t)code[MB_P_LOGA][(size_t)g];
    uint32_t od = 0, ci = 0;
    for (int p : dust_free) {
      od += (uint32_t)code[p][(size_t)g] * h->oradix[p];
      ci += (uint32_t)code[p][(size_t)g] * h->iradix[p];
    }
    rowcode[(size_t)g] = ((cA * (uint32_t)h->nOd + od) << kAgeShift) | ci;
  }
  // look-up rows behind every table row, and where a look-up row lives while K2 builds its tables:
  // a parameter that is alone in its table is evaluated in place, the others in a scratch table
  std::vector<uint16_t> rowsrc((size_t)(h->nO + h->nI) * 4, (uint16_t)0xffff);
  std::vector<uint16_t> lutslot((size_t)std::max(1, clsoff), (uint16_t)0), rowslot;
  if (!elementwise) {
    const bool age_free = model->kind[MB_P_LOGA] != MB_FIXED;
    int n_outer_params = age_free ? 1 : 0, n_inner_params = 0;
    for (int p : dust_free) (((inner_mask >> (p - MB_P_AV)) & 1) ? n_inner_params : n_outer_params)++;
    for (int r = 0; r < h->nO; ++r) {
      int k = 0;
      if (age_free) rowsrc[(size_t)r * 4 + k++] = (uint16_t)(md.clsoff[MB_P_LOGA] + r / h->nOd);
      for (int p : dust_free)
        if (!((inner_mask >> (p - MB_P_AV)) & 1))
          rowsrc[(size_t)r * 4 + k++] = (uint16_t)(md.clsoff[p] + ((r % h->nOd) / h->oradix[p]) % md.ncls[p]);
    }
    for (int r = 0; r < h->nI; ++r) {
      int k = 0;
      for (int p : dust_free)
        if ((inner_mask >> (p - MB_P_AV)) & 1)
          rowsrc[(size_t)(h->nO + r) * 4 + k++] = (uint16_t)(md.clsoff[p] + (r / h->iradix[p]) % md.ncls[p]);
    }
    int nscratch = 0;
    for (int p = 0; p < MB_NPARAM; ++p) {
      if (model->kind[p] == MB_FIXED) continue;
      const bool inner = p != MB_P_LOGA && ((inner_mask >> (p - MB_P_AV)) & 1);
      const bool alone = inner ? n_inner_params == 1 : n_outer_params == 1;
      if (p == MB_P_LOGA) h->age_slot = alone ? -1 : nscratch;  // -1: the age rows are the outer table itself
      for (int c = 0; c < md.ncls[p]; ++c)
        lutslot[(size_t)md.clsoff[p] + c] =
            alone ? (uint16_t)(0x8000u | (unsigned)((inner ? h->nO : 0) + c)) : (uint16_t)nscratch++;
    }
    h->nscratch = nscratch;
    if (h->nO + h->nI >= 0x7ff0) {
      // (e.g. a small grid whose every row has its own A(V) and R(V): few enough classes per column to be
      // coded, but their product is as large as the grid squared)
      fail("more than 32751 factor-table rows");
      return (o.mode == MB_MODE_AUTO && n_pixels == 0) ? kRetryElementwise : 1;
    }
    // K2's own table build reads the sources of a product row as slots directly
    rowslot.assign(rowsrc.size(), (uint16_t)0xffff);
    for (int r = 0; r < h->nO + h->nI; ++r) {
      const uint16_t* src = &rowsrc[(size_t)r * 4];
      if (src[0] != 0xffff && src[1] == 0xffff && lutslot[src[0]] == (uint16_t)(0x8000u | (unsigned)r)) {
        rowslot[(size_t)r * 4] = 0xfffe;
        continue;
      }
      for (int k = 0; k < 4; ++k)
        if (src[k] != 0xffff) rowslot[(size_t)r * 4 + k] = lutslot[src[k]];
    }
  }

  // ---- N-stars sufficient statistics (helpers.py:133-162) ----------------------------------------
  std::vector<double> H, coef, bin_age;
  std::vector<int32_t> bin_agecls, binrows;
  std::vector<int64_t> binoff, tile_start, bin_tile;
  if (h->poisson && elementwise) {
    // ELEMENTWISE: no statistics can be folded ahead of time; K1b sums physmod * grid_weight
    // (* completeness) over each bin's rows, listed here in bin order
    if (!grid->logA || !grid->Z) return fail("logA and Z columns are required for the N-stars term");
    if (mm->n_ages < 1 || mm->n_mets < 1 || !mm->ages || !mm->mets || !mm->massmult)
      return fail("invalid mass-multiplier description");
    const int nAg = mm->n_ages, nZ = mm->n_mets;
    h->nbins = nAg * nZ;
    if (h->nbins > 2048) return fail("N-stars term: more than 2048 (age, Z) bins (%d)", h->nbins);
    std::vector<int32_t> rowbin((size_t)G);
    binoff.assign((size_t)h->nbins + 1, 0);
    for (int64_t g = 0; g < G; ++g) {
      const double* ia = std::lower_bound(mm->ages, mm->ages + nAg, grid->logA[g]);
      const double* iz = std::lower_bound(mm->mets, mm->mets + nZ, grid->Z[g]);
      if (ia == mm->ages + nAg || *ia != grid->logA[g] || iz == mm->mets + nZ || *iz != grid->Z[g])
        return fail("grid row %lld has (logA, Z) outside the mass-multiplier bins", (long long)g);
      const int32_t b = (int32_t)((ia - mm->ages) * nZ + (iz - mm->mets));
      rowbin[(size_t)g] = b;
      ++binoff[(size_t)b + 1];
    }
    for (int b = 0; b < h->nbins; ++b) binoff[(size_t)b + 1] += binoff[(size_t)b];
    binrows.resize((size_t)G);
    std::vector<int64_t> fill(binoff.begin(), binoff.end() - 1);
    for (int64_t g = 0; g < G; ++g) binrows[(size_t)fill[(size_t)rowbin[(size_t)g]]++] = (int32_t)g;
    // K1e tiles: at most kK1eRows consecutive positions of the bin-ordered list, inside one bin
    bin_tile.assign(1, 0);
    for (int b = 0; b < h->nbins; ++b) {
      for (int64_t i = binoff[(size_t)b]; i < binoff[(size_t)b + 1]; i += kK1eRows) tile_start.push_back(i);
      bin_tile.push_back((int64_t)tile_start.size());
    }
    tile_start.push_back(G);
    h->ntiles = (int64_t)tile_start.size() - 1;
    coef.resize((size_t)h->nbins);
    bin_age.resize((size_t)h->nbins);
    const double metprior = 1.0 / (double)nZ;
    for (int i = 0; i < nAg; ++i)
      for (int j = 0; j < nZ; ++j) {
        coef[(size_t)i * nZ + j] = metprior * mm->massmult[(size_t)i * nZ + j] / mm->avemass;
        bin_age[(size_t)i * nZ + j] = mm->ages[i];
      }
    bin_agecls.resize((size_t)h->nbins);   // K1e leaves the age factor per BIN: row of `fage` = the bin itself
    for (int b = 0; b < h->nbins; ++b) bin_agecls[(size_t)b] = b;
  } else if (h->poisson) {
    if (!grid->logA || !grid->Z) return fail("logA and Z columns are required for the N-stars term");
    if (mm->n_ages < 1 || mm->n_mets < 1 || !mm->ages || !mm->mets || !mm->massmult)
      return fail("invalid mass-multiplier description");
    const int nAg = mm->n_ages, nZ = mm->n_mets;
    h->nbins = nAg * nZ;
    if ((int64_t)h->nbins * h->nB > (1LL << 26) || h->nbins > 2048)
      return fail("N-stars statistics too large (%d bins x %d dust classes)", h->nbins, h->nB);
    H.assign((size_t)h->nbins * h->nB * 2, 0.0);
    std::vector<int32_t> age_first_row((size_t)nAg, -1);
    for (int64_t g = 0; g < G; ++g) {
      const double* ia = std::lower_bound(mm->ages, mm->ages + nAg, grid->logA[g]);
      const double* iz = std::lower_bound(mm->mets, mm->mets + nZ, grid->Z[g]);
      if (ia == mm->ages + nAg || *ia != grid->logA[g] || iz == mm->mets + nZ || *iz != grid->Z[g])
        return fail("grid row %lld has (logA, Z) outside the mass-multiplier bins", (long long)g);
      int i = (int)(ia - mm->ages), j = (int)(iz - mm->mets);
      if (age_first_row[(size_t)i] < 0) age_first_row[(size_t)i] = (int32_t)g;
      size_t b = (size_t)i * nZ + j;
      const uint32_t rc = rowcode[(size_t)g];
      // dust class in table order: outer-dust digit * nI + inner class
      const size_t cD = (size_t)(((rc >> kAgeShift) & kAgeMask) % (uint32_t)h->nOd) * h->nI + (rc & kDustMask);
      double gw = grid->grid_weight[g];
      // K2 reads a bin's classes contiguously: Ht[bin][cD]; the pixel kernel (lanes = bins): Ht[cD][bin]
      const size_t at = n_pixels > 0 ? (cD * h->nbins + b) * 2 : (b * (size_t)h->nB + cD) * 2;
      H[at] += gw;
      H[at + 1] += gw * grid->completeness[g];
    }
    coef.resize((size_t)h->nbins);
    bin_agecls.resize((size_t)h->nbins);
    const double metprior = 1.0 / (double)nZ;
    for (int i = 0; i < nAg; ++i)
      for (int j = 0; j < nZ; ++j) {
        coef[(size_t)i * nZ + j] = metprior * mm->massmult[(size_t)i * nZ + j] / mm->avemass;
        // an age no grid row carries has empty bins (S = 0, skipped); any class will do
        int32_t r = age_first_row[(size_t)i];
        bin_agecls[(size_t)i * nZ + j] = r < 0 ? 0 : (int32_t)(((rowcode[(size_t)r] >> kAgeShift) & kAgeMask) / (uint32_t)h->nOd);
      }
  }

  // ---- entries: fold, code, sort by age class, (merge), flag -------------------------------------
  const int64_t K = stars->K, S = stars->S;
  int64_t s0 = stars->star_begin, s1 = stars->star_end;
  if (n_pixels > 0) {
    if (!pix_off) return fail("pixel star offsets are NULL");
    for (int64_t p = 0; p < n_pixels; ++p)
      if (pix_off[p] > pix_off[p + 1]) return fail("pixel star offsets must be non-decreasing");
    s0 = pix_off[0];
    s1 = pix_off[n_pixels];
  }
  if (K < 0 || S < 0 || s0 < 0 || s1 > S || s0 > s1) return fail("invalid star range [%lld,%lld) of %lld", (long long)s0, (long long)s1, (long long)S);
  if (K > 0 && S > 0 && (!stars->indxs || !stars->vals)) return fail("indxs/vals are NULL");
  h->S_local = s1 - s0;
  h->S_total = stars->n_stars_total > 0 ? stars->n_stars_total : S;
  // 16-bit code words where the class counts allow: 4 + 10 or 6 + 8 bits of (outer, inner) class
  h->ob = 0;
  if (mode == MB_MODE_FACTOR && h->table_bytes == 8 && n_pixels == 0 && !getenv("MB_CODES32")) {
    if (h->nO <= 16 && h->nI <= 1024) h->ob = 4;
    else if (h->nO <= 64 && h->nI <= 256) h->ob = 6;
  }
  h->c16 = h->ob > 0;
  const bool merge = o.merge_duplicates != 0;
  std::vector<int64_t> star_end_pos;  // entry count after each non-empty star
  std::vector<int64_t> pix_ent, pix_se;  // pixel batches: entry / star-end bookkeeping per pixel
  if (o.keep_indices) {  // parity hook: the grid rows that are dereferenced, in the reference's order
    h->kept_offs.assign(1, 0);
    for (int64_t s = s0; s < s1; ++s) {
      for (int64_t kk = 0; kk < K; ++kk) {
        if (!std::isfinite(stars->vals[kk * S + s])) continue;
        int64_t r = stars->indxs[kk * S + s];
        h->kept_rows.push_back(r < 0 ? r + G : r);
      }
      h->kept_offs.push_back((int64_t)h->kept_rows.size());
    }
  }
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  for (auto& e : h->ev) CK(cudaEventCreate(&e));
  // The build runs on the device (mb_build_entries); the host loop below is the same algorithm, kept
  // for stars with more than 128 sparse points and as an A/B switch (MB_HOST_BUILD=1) -- the two
  // produce bit-identical entry streams (tests/test_gpu_split_tables.py::test_device_build_equals_host_build)
  const bool host_build = K > 128 || (getenv("MB_HOST_BUILD") && atoi(getenv("MB_HOST_BUILD")));
  if (!host_build) {
    if (build_entries_device(h, grid, stars, s0, s1, elementwise ? nullptr : rowcode.data(), mode, merge,
                             n_pixels, pix_off, star_end_pos, pix_ent, pix_se)) return 1;
  } else {
    std::vector<Entry> entries;
    std::vector<uint32_t> codes;
    entries.reserve((size_t)((s1 - s0) * K));
    codes.reserve((size_t)((s1 - s0) * K));
    std::vector<Entry> cur;
    int64_t next_pix = 0;
    for (int64_t s = s0; s < s1; ++s) {
      while (next_pix < n_pixels && pix_off[next_pix] == s) {
        pix_ent.push_back((int64_t)entries.size());
        pix_se.push_back((int64_t)star_end_pos.size());
        ++next_pix;
      }
      cur.clear();
      for (int64_t kk = 0; kk < K; ++kk) {
        double v = stars->vals[kk * S + s];
        if (!std::isfinite(v)) continue;  // :196 -- the index of a masked entry is never read
        int64_t r = stars->indxs[kk * S + s];
        if (r < 0) r += G;  // numpy negative indexing
        if (r < 0 || r >= G)
          return fail("index %lld is out of bounds for axis 0 with size %lld (star %lld)",
                      (long long)stars->indxs[kk * S + s], (long long)G, (long long)s);
        Entry e;
        e.a = v * grid->prior_weight[r] * grid->grid_weight[r] * grid->completeness[r];
        const uint32_t rc = elementwise ? 0u : rowcode[(size_t)r];
        // (TABLE modes always use the classic split: outer class = age class, inner class = dust class)
        const uint32_t cO = (rc >> kAgeShift) & kAgeMask, cI = rc & kDustMask;
        e.meta = cO;
        e.idx = mode == MB_MODE_FACTOR ? cI : mode == MB_MODE_TABLE_UNIQUE ? cO * (uint32_t)h->nB + cI : (uint32_t)r;
        cur.push_back(e);
      }
      h->n_entries_raw += (int64_t)cur.size();
      if (cur.empty()) continue;  // sums to 0.0 -> skipped by the reference too (:218-221)
      // runs of equal outer class (FACTOR reloads the outer factor once per run); ties by table row
      std::stable_sort(cur.begin(), cur.end(), [](const Entry& x, const Entry& y) {
        return x.meta != y.meta ? x.meta < y.meta : x.idx < y.idx;
      });
      size_t first = entries.size();
      for (const Entry& e : cur) {
        if (merge && entries.size() > first) {
          Entry& last = entries.back();
          if (last.idx == e.idx && last.meta == e.meta) { last.a += e.a; continue; }
        }
        entries.push_back(e);
      }
      // device code words: FACTOR [31] new run, [30] new star, [29:20] outer class, [19:0] inner class;
      // TABLE modes read the full product from the table: [31] new star, [30:0] table row
      uint32_t prevA = 0xffffffffu;
      for (size_t i = first; i < entries.size(); ++i) {
        const uint32_t cA = entries[i].meta;
        uint32_t code;
        if (mode == MB_MODE_FACTOR) {
          code = (cA << kAgeShift) | entries[i].idx;
          if (i == first) code |= kNewRun | kNewStar;
          else if (cA != prevA) code |= kNewRun;
          if (code & kNewRun) ++h->n_runs;
        } else {
          code = entries[i].idx;
          if (i == first) code |= kNewRun;
        }
        codes.push_back(code);
        prevA = cA;
      }
      star_end_pos.push_back((int64_t)entries.size());
    }
    h->n_entries = (int64_t)entries.size();
    while (next_pix <= n_pixels && n_pixels > 0) {  // trailing empty pixels + the end sentinel
      pix_ent.push_back((int64_t)entries.size());
      pix_se.push_back((int64_t)star_end_pos.size());
      ++next_pix;
    }

    // upload
    // (padded: the ring copies whole chunks, so it reads up to a chunk past the last entry)
    std::vector<double> wts(entries.size() + kEntryPad, 0.0);
    for (size_t i = 0; i < entries.size(); ++i) wts[i] = entries[i].a;
    if (dev_upload(h, &h->d_ea, wts)) return 1;
    codes.resize(codes.size() + kEntryPad, 0u);
    if (h->c16) {
      std::vector<uint16_t> c16(codes.size(), (uint16_t)0);
      for (size_t i = 0; i < codes.size(); ++i) {
        const uint32_t c = codes[i];
        c16[i] = (uint16_t)(((c & kNewRun) ? 0x8000u : 0u) | ((c & kNewStar) ? 0x4000u : 0u) |
                            (((c >> kAgeShift) & kAgeMask) << (14 - h->ob)) | (c & kDustMask));
      }
      if (dev_upload(h, &h->d_ec16, c16)) return 1;
    } else if (dev_upload(h, &h->d_ec, codes)) return 1;
  }

  // ---- segments: star aligned, one table per block shape ---------------------------------------------
  // measured on B200 (config 3 and its 1/2, 1/8 shards): 1 segment per resident warp beats 2, 4, 8
  // by 3-4 % -- per-segment costs (pipeline restart, one log) outweigh what dynamic balancing
  // recovers -- so the segment count must match the warps the launch really has: one table for
  // the 768-thread kernels (up to 64 walkers per pass) and one for the 512-thread kernel (128).
  h->spw = o.segments_per_warp > 0 ? o.segments_per_warp : 1;
  // ... optionally plus a tail pool: the last tail_percent of the entries cut into small segments
  // that warps claim dynamically after their own big one.  Measured on B200 (4-16 %, config 3 and
  // its 1/8 shard): no gain over none, so the default is none.
  h->tail_frac = o.tail_percent > 0 ? std::min(50, o.tail_percent) / 100.0 : 0.0;
  if (const char* pt = getenv("MB_PEER_TIMEOUT_MS")) h->peer_timeout_ns = (unsigned long long)std::max(1.0, atof(pt)) * 1000000ull;
  h->star_end_pos = std::move(star_end_pos);
  std::vector<int64_t> seg[2];
  for (int t = 0; t < 2; ++t) {
    build_segments(h, t, seg[t]);
    h->nseg[t] = (int)seg[t].size() - 1;
  }

  // ---- upload ----------------------------------------------------------------------------------------
  std::vector<double> clsx_all;
  for (int p = 0; p < MB_NPARAM; ++p) clsx_all.insert(clsx_all.end(), h->clsx[p].begin(), h->clsx[p].end());
  if (dev_upload(h, &h->d_clsx, clsx_all)) return 1;
  if (dev_upload(h, &h->d_seg[0], seg[0])) return 1;
  if (dev_upload(h, &h->d_seg[1], seg[1])) return 1;
  if (mode == MB_MODE_TABLE_GRID && dev_upload(h, &h->d_rowcode, rowcode)) return 1;
  if (!elementwise) {
    if (dev_upload(h, &h->d_rowsrc, rowsrc)) return 1;
    if (dev_upload(h, &h->d_lutslot, lutslot)) return 1;
    if (dev_upload(h, &h->d_rowslot, rowslot)) return 1;
  }
  if (h->poisson && dev_alloc(h, &h->d_Pb, (int64_t)h->nbins * 2 * 128)) return 1;
  if (elementwise) {
    for (int p = 0; p < MB_NPARAM; ++p) {
      if (model->kind[p] == MB_FIXED) continue;
      CK(cudaMalloc((void**)&h->d_col[p], sizeof(double) * (size_t)G));
      h->device_bytes += 8 * G;
      CK(cudaMemcpy(h->d_col[p], grid->col[p], sizeof(double) * (size_t)G, cudaMemcpyHostToDevice));
    }
    if (h->poisson) {
      CK(cudaMalloc((void**)&h->d_gw, sizeof(double) * (size_t)G));
      CK(cudaMalloc((void**)&h->d_compl, sizeof(double) * (size_t)G));
      h->device_bytes += 16 * G;
      CK(cudaMemcpy(h->d_gw, grid->grid_weight, sizeof(double) * (size_t)G, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(h->d_compl, grid->completeness, sizeof(double) * (size_t)G, cudaMemcpyHostToDevice));
      if (dev_upload(h, &h->d_binrows, binrows)) return 1;
      if (dev_upload(h, &h->d_tile_start, tile_start)) return 1;
      if (dev_upload(h, &h->d_bin_tile, bin_tile)) return 1;
      if (dev_alloc(h, &h->d_partial, h->ntiles * 128 * 2)) return 1;
      if (dev_upload(h, &h->d_bin_age, bin_age)) return 1;
      if (dev_upload(h, &h->d_coef, coef)) return 1;
      if (dev_upload(h, &h->d_bin_agecls, bin_agecls)) return 1;
      if (dev_alloc(h, &h->d_Fage, (int64_t)h->nbins * 128)) return 1;
      if (dev_alloc(h, &h->d_SC, (int64_t)h->nbins * 128 * 2)) return 1;
    }
  } else if (h->poisson) {
    if (dev_upload(h, &h->d_H, H)) return 1;
    if (dev_upload(h, &h->d_coef, coef)) return 1;
    if (dev_upload(h, &h->d_bin_agecls, bin_agecls)) return 1;
  }
  if (dev_alloc(h, &h->d_F, (int64_t)(h->nO + h->nI + h->nA) * 128)) return 1;
  if (dev_alloc(h, &h->d_counter, kCtrWords)) return 1;
  CK(cudaMemset(h->d_counter, 0, sizeof(unsigned int) * kCtrWords));
  if (dev_alloc(h, &h->d_gacc, kAccRows * 128)) return 1;
  CK(cudaMemset(h->d_gacc, 0, sizeof(unsigned long long) * kAccRows * 128));
  h->max_blocks = h->sm_count * 4;
  CK(cudaMalloc((void**)&h->d_xbuf, kXBufBytes));
  CK(cudaMemset(h->d_xbuf, 0, kXBufBytes));
  h->device_bytes += (int64_t)kXBufBytes;

  // opt in to large dynamic shared memory for every K2 instantiation we may launch
  {
    int big = h->max_smem;
#define MB_OPTIN(NW, MODE) CK(cudaFuncSetAttribute(mb_k2_stars<NW, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, big))
    MB_OPTIN(1, kModeFactor); MB_OPTIN(2, kModeFactor); MB_OPTIN(4, kModeFactor);
    CK(cudaFuncSetAttribute(mb_k2_stars<1, kModeFactor, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CK(cudaFuncSetAttribute(mb_k2_stars<2, kModeFactor, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CK(cudaFuncSetAttribute(mb_k2_stars<4, kModeFactor, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CK((cudaFuncSetAttribute(mb_k2_stars<1, kModeFactor, double, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)));
    CK((cudaFuncSetAttribute(mb_k2_stars<2, kModeFactor, double, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)));
    CK((cudaFuncSetAttribute(mb_k2_stars<4, kModeFactor, double, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)));
    CK((cudaFuncSetAttribute(mb_k2_stars<1, kModeFactor, double, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)));
    CK((cudaFuncSetAttribute(mb_k2_stars<2, kModeFactor, double, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)));
    CK((cudaFuncSetAttribute(mb_k2_stars<4, kModeFactor, double, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)));
    MB_OPTIN(1, kModeTable); MB_OPTIN(2, kModeTable); MB_OPTIN(4, kModeTable);
    MB_OPTIN(1, kModeTableGrid); MB_OPTIN(2, kModeTableGrid); MB_OPTIN(4, kModeTableGrid);
#undef MB_OPTIN
  }
  if (elementwise) {
    const size_t need = k1e_smem_bytes(128, h->ndim);
    if (need > (size_t)h->max_smem) return fail("ELEMENTWISE: %d free variables do not fit shared memory", h->ndim);
    CK(cudaFuncSetAttribute(mb_k1_elementwise<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    CK(cudaFuncSetAttribute(mb_k1_elementwise<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
  }
  if (n_pixels > 0) {
    if (mode != MB_MODE_FACTOR) return fail("pixel batches need the factor tables in shared memory (FACTOR mode)");
    h->P = n_pixels;
    std::vector<int64_t> pseg((size_t)n_pixels * (kPixWarps + 1));
    std::vector<int32_t> pns((size_t)n_pixels);
    for (int64_t p = 0; p < n_pixels; ++p) {
      const int64_t e0 = pix_ent[(size_t)p], e1 = pix_ent[(size_t)p + 1];
      const std::vector<int64_t>& star_end_pos = h->star_end_pos;
      size_t j = (size_t)pix_se[(size_t)p];
      const size_t j1 = (size_t)pix_se[(size_t)p + 1];
      int64_t* row = &pseg[(size_t)p * (kPixWarps + 1)];
      row[0] = e0;
      for (int i = 1; i <= kPixWarps; ++i) {  // boundary i: first star end at or after the i/12 point
        const int64_t goal = e0 + ((e1 - e0) * i + kPixWarps - 1) / kPixWarps;
        while (j < j1 && star_end_pos[j] < goal) ++j;
        row[i] = i == kPixWarps ? e1 : (j < j1 ? star_end_pos[j] : e1);
      }
      pns[(size_t)p] = (int32_t)(pix_off[p + 1] - pix_off[p]);
    }
    if (dev_upload(h, &h->d_pix_seg, pseg)) return 1;
    if (dev_upload(h, &h->d_pix_nstars, pns)) return 1;
    CK(cudaFuncSetAttribute(mb_k2_pixels<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
    CK(cudaFuncSetAttribute(mb_k2_pixels<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
    CK(cudaFuncSetAttribute(mb_k2_pixels<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
  }
  guard.ok = true;
  *out = h;
  return 0;
}

extern "C" int mb_create(const mb_grid_desc* grid, const mb_stars_desc* stars,
                         const mb_model_desc* model, const mb_massmult_desc* mm,
                         const mb_options* opts, mb_handle** out) {
  int rc = create_impl(grid, stars, model, mm, opts, 0, nullptr, out);
  if (rc == kRetryElementwise) {
    mb_options o2{};
    if (opts) o2 = *opts;
    o2.mode = MB_MODE_ELEMENTWISE;
    rc = create_impl(grid, stars, model, mm, &o2, 0, nullptr, out);
  }
  return rc ? 1 : 0;
}

extern "C" int mb_create_pixels(const mb_grid_desc* grid, const mb_stars_desc* stars,
                                const mb_model_desc* model, const mb_massmult_desc* mm,
                                const mb_options* opts, int64_t n_pixels, const int64_t* star_offsets,
                                mb_handle** out) {
  if (n_pixels <= 0) return fail("mb_create_pixels: need at least one pixel");
  return create_impl(grid, stars, model, mm, opts, n_pixels, star_offsets, out);
}

// ---- pixel batches: P independent ensembles in one launch -------------------------------------------
static size_t pix_front_bytes(int nbins, int Wpad) {
  size_t stage = (size_t)kPixWarps * kStageBytes;
  size_t red = nbins > 0 ? (size_t)(Wpad / 8) * ((nbins + 31) / 32) * 16 * sizeof(double) : 0;
  return (std::max(stage, red) + 15) & ~(size_t)15;
}
static size_t pix_smem_bytes(int nA, int nB, int nbins, int Wpad) {
  return pix_front_bytes(nbins, Wpad) + (size_t)(nA + nB) * Wpad * 8 + (size_t)Wpad * 8 + (size_t)kAccRows * Wpad * 8;
}

static int eval_pixels_device(mb_handle* h, const double* d_phi, int W, int ndim, double* d_out, cudaStream_t st,
                              const double* const d_tab[MB_NPARAM] = nullptr, const double* d_box = nullptr) {
  if (h->P <= 0) return fail("handle was not created with mb_create_pixels");
  if (ndim != h->ndim) return fail("phi has %d columns, the model has %d free variables", ndim, h->ndim);
  if (W < 1 || W > 128) return fail("pixel batches take 1..128 walkers per pixel (got %d)", W);
  for (int p = 0; p < MB_NPARAM; ++p)
    if (h->model.kind[p] == MB_TABULATED && !(d_tab && d_tab[p]))
      return fail("parameter %d is tabulated but no table was passed (mb_lnlike_pixels_tabulated)", p);
  CK(cudaSetDevice(h->device));
  int NW = W <= 32 ? 1 : W <= 64 ? 2 : 4;
  const int Wpad = 32 * NW;
  const size_t smem = pix_smem_bytes(h->nA, h->nB, h->nbins, Wpad);
  if (smem > (size_t)h->max_smem) return fail("pixel kernel needs %zu bytes of shared memory", smem);
  const int64_t needF = h->P * (int64_t)(h->nA + h->nB + h->nA) * Wpad;  // K0's layout: age, dust, plain age rows
  if (needF > h->cap_Fpix) {
    CK(cudaStreamSynchronize(st));
    if (h->d_Fpix) { CK(cudaFree(h->d_Fpix)); h->device_bytes -= h->cap_Fpix * 8; }
    h->d_Fpix = nullptr; h->cap_Fpix = 0;
    if (dev_alloc(h, &h->d_Fpix, needF)) return 1;
    h->cap_Fpix = needF;
  }
  const bool prof = h->profiling != 0;
  if (prof) CK(cudaEventRecord(h->ev[0], st));
  K0Args k0{};
  k0.model = h->mdev; k0.clsx = h->d_clsx; k0.phi = d_phi; k0.F = h->d_Fpix; k0.box = d_box;
  for (int p = 0; p < MB_NPARAM; ++p) k0.tab[p] = d_tab ? d_tab[p] : nullptr;
  k0.rowsrc = h->d_rowsrc; k0.nrows = h->nA + h->nB; k0.nAge = h->nA;
  k0.W = W; k0.Wpad = Wpad; k0.ndim = ndim; k0.ncls_total = h->ncls_total;
  mb_k0_factors<<<(unsigned)(h->P * Wpad), kK0Threads, sizeof(double) * std::max(1, h->ncls_total), st>>>(k0);
  if (prof) { CK(cudaEventRecord(h->ev[1], st)); CK(cudaEventRecord(h->ev[2], st)); }
  KPixArgs ka{};
  ka.ea = h->d_ea; ka.ec = h->d_ec; ka.pix_seg = h->d_pix_seg; ka.F = h->d_Fpix; ka.pix_nstars = h->d_pix_nstars;
  ka.out = d_out; ka.counters = h->d_counter;
  ka.pois.Ht = h->d_H; ka.pois.bin_agecls = h->d_bin_agecls; ka.pois.coef = h->d_coef;
  ka.pois.nbins = h->poisson ? h->nbins : 0; ka.pois.n_stars = 0.0;
  ka.P = (int)h->P; ka.nA = h->nA; ka.nB = h->nB; ka.Wpad = Wpad; ka.W = W;
  ka.tab_off = (int32_t)pix_front_bytes(h->nbins, Wpad);
  int per_sm = 0;
  switch (NW) {
    case 1: CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mb_k2_pixels<1>, kPixThreads, smem)); break;
    case 2: CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mb_k2_pixels<2>, kPixThreads, smem)); break;
    default: CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mb_k2_pixels<4>, kPixThreads, smem)); break;
  }
  if (per_sm < 1) return fail("pixel kernel cannot be resident");
  const int grid = (int)std::min<int64_t>(h->P, (int64_t)h->sm_count * per_sm);
  switch (NW) {
    case 1: mb_k2_pixels<1><<<grid, kPixThreads, smem, st>>>(ka); break;
    case 2: mb_k2_pixels<2><<<grid, kPixThreads, smem, st>>>(ka); break;
    default: mb_k2_pixels<4><<<grid, kPixThreads, smem, st>>>(ka); break;
  }
  if (prof) {
    CK(cudaEventRecord(h->ev[3], st));
    CK(cudaEventSynchronize(h->ev[3]));
    for (int i = 0; i < MB_NKERNELS; ++i) CK(cudaEventElapsedTime(&h->last_ms[i], h->ev[i], h->ev[i + 1]));
  }
  CK(cudaGetLastError());
  h->last_wp = W;
  h->last_launches = 2;
  return 0;
}

extern "C" int mb_lnlike_pixels_device(mb_handle* h, const double* d_phi, int32_t W, int32_t ndim,
                                       double* d_out, void* stream) {
  if (!h || !d_phi || !d_out) return fail("mb_lnlike_pixels_device: null argument");
  return eval_pixels_device(h, d_phi, W, ndim, d_out, (cudaStream_t)stream);
}

extern "C" int64_t mb_pixel_count(const mb_handle* h) { return h ? h->P : 0; }

extern "C" int mb_get_info(const mb_handle* h, mb_info* out) {
  if (!h || !out) return fail("mb_get_info: null argument");
  memset(out, 0, sizeof *out);
  out->abi_version = MB_ABI_VERSION;
  out->device = h->device;
  out->mode = h->mode;
  out->ndim = h->ndim;
  out->poisson = h->poisson;
  for (int p = 0; p < MB_NPARAM; ++p) out->n_classes[p] = h->mdev.ncls[p];
  out->n_age_classes = h->nA;
  out->n_dust_classes = h->nB;
  out->n_bins = h->nbins;
  out->sm_count = h->sm_count;
  out->walkers_per_pass = h->last_wp;
  out->launches_last_eval = h->last_launches;
  out->G = h->G;
  out->n_stars_local = h->S_local;
  out->n_stars_total = h->S_total;
  out->n_entries = h->n_entries;
  out->n_entries_raw = h->n_entries_raw;
  out->device_bytes = h->device_bytes;
  out->code_bytes = h->c16 ? 2 : 4;
  out->n_outer_rows = h->nO;
  out->n_inner_rows = h->nI;
  out->split_inner_mask = h->inner_mask;
  out->n_runs = h->n_runs;
  return 0;
}

extern "C" int mb_class_values(const mb_handle* h, int32_t p, double* out, int32_t cap) {
  if (!h || p < 0 || p >= MB_NPARAM) { fail("mb_class_values: bad argument"); return -1; }
  int n = (int)h->clsx[p].size();
  for (int i = 0; i < n && i < cap; ++i) out[i] = h->clsx[p][(size_t)i];
  return n;
}

extern "C" int mb_gathered_indices(const mb_handle* h, int64_t* rows, int64_t cap, int64_t* offs) {
  if (!h) return fail("null handle");
  if (h->kept_offs.empty()) return fail("handle was created without keep_indices");
  if ((int64_t)h->kept_rows.size() > cap) return fail("rows buffer too small (%lld needed)", (long long)h->kept_rows.size());
  if ((h->mode == MB_MODE_TABLE_GRID || h->mode == MB_MODE_ELEMENTWISE) && h->n_entries == h->n_entries_raw) {
    // TABLE_GRID gathers by grid row: report what the device actually holds (round trip)
    std::vector<uint32_t> back((size_t)h->n_entries);
    if (h->n_entries) CK(cudaMemcpy(back.data(), h->d_ec, sizeof(uint32_t) * back.size(), cudaMemcpyDeviceToHost));
    size_t i = 0, nout = 0;
    for (int64_t s = 0; s < h->S_local; ++s) {
      offs[s] = (int64_t)nout;
      int64_t cnt = h->kept_offs[(size_t)s + 1] - h->kept_offs[(size_t)s];
      for (int64_t c = 0; c < cnt; ++c) rows[nout++] = (int64_t)(back[i++] & kRowMask);
    }
    offs[h->S_local] = (int64_t)nout;
    return 0;
  }
  // class-coded modes dereference the grid only at create time: the rows that were read then
  memcpy(rows, h->kept_rows.data(), sizeof(int64_t) * h->kept_rows.size());
  memcpy(offs, h->kept_offs.data(), sizeof(int64_t) * h->kept_offs.size());
  return 0;
}

extern "C" int mb_device_entries(const mb_handle* h, int64_t cap, int32_t* cls, double* weight, int64_t* offs) {
  if (!h || !cls || !weight || !offs) return fail("mb_device_entries: null argument");
  if (h->kept_offs.empty()) return fail("handle was created without keep_indices");
  if (h->mode != MB_MODE_FACTOR && h->mode != MB_MODE_TABLE_UNIQUE)
    return fail("mb_device_entries: the device holds class codes in the FACTOR and TABLE_UNIQUE modes only");
  if (h->n_entries > cap) return fail("entry buffers too small (%lld needed)", (long long)h->n_entries);
  CK(cudaSetDevice(h->device));
  const size_t n = (size_t)h->n_entries;
  std::vector<uint32_t> code(n);
  if (n) {
    CK(cudaMemcpy(weight, h->d_ea, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (h->c16) {
      std::vector<uint16_t> c16(n);
      CK(cudaMemcpy(c16.data(), h->d_ec16, sizeof(uint16_t) * n, cudaMemcpyDeviceToHost));
      const int ib = 14 - h->ob;
      for (size_t i = 0; i < n; ++i) {  // back to the 32-bit layout
        const uint32_t c = c16[i];
        code[i] = ((c & 0x8000u) ? kNewRun : 0u) | ((c & 0x4000u) ? kNewStar : 0u) |
                  (((c >> ib) & ((1u << h->ob) - 1u)) << kAgeShift) | (c & ((1u << ib) - 1u));
      }
    } else {
      CK(cudaMemcpy(code.data(), h->d_ec, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
    }
  }
  const ModelDev& md = h->mdev;
  size_t e = 0;
  for (int64_t s = 0; s < h->S_local; ++s) {
    offs[s] = (int64_t)e;
    if (h->kept_offs[(size_t)s + 1] == h->kept_offs[(size_t)s]) continue;  // no finite entry: nothing on the device
    // this star's entries: from its new-star flag up to the next one
    bool first = true;
    for (; e < n; ++e) {
      const uint32_t c = code[e];
      const bool new_star = h->mode == MB_MODE_FACTOR ? (c & kNewStar) != 0 : (c & kNewRun) != 0;
      if (new_star && !first) break;
      if (first && !new_star) return fail("mb_device_entries: star %lld does not start with a new-star flag", (long long)s);
      first = false;
      uint32_t cA, od, ci;
      if (h->mode == MB_MODE_FACTOR) {
        const uint32_t cO = (c >> kAgeShift) & kAgeMask;
        cA = cO / (uint32_t)h->nOd; od = cO % (uint32_t)h->nOd; ci = c & kDustMask;
      } else {
        const uint32_t u = c & kRowMask;
        cA = u / (uint32_t)h->nB; od = 0; ci = u % (uint32_t)h->nB;
      }
      int32_t* out = cls + e * MB_NPARAM;
      out[MB_P_LOGA] = (int32_t)cA;
      for (int p = MB_P_AV; p < MB_NPARAM; ++p) {
        out[p] = 0;
        if (md.kind[p] == MB_FIXED) continue;
        out[p] = (h->inner_mask >> (p - MB_P_AV)) & 1 ? (int32_t)((ci / h->iradix[p]) % (uint32_t)md.ncls[p])
                                                      : (int32_t)((od / h->oradix[p]) % (uint32_t)md.ncls[p]);
      }
    }
    if (first) return fail("mb_device_entries: star %lld has no entries on the device", (long long)s);
  }
  offs[h->S_local] = (int64_t)e;
  if (e != n) return fail("mb_device_entries: %lld entries left over after the last star", (long long)(n - e));
  return 0;
}

extern "C" int mb_set_profiling(mb_handle* h, int32_t on) {
  if (!h) return fail("null handle");
  h->profiling = on;
  if (on && !h->d_timeline) {
    CK(cudaSetDevice(h->device));
    if (dev_alloc(h, &h->d_timeline, (int64_t)h->max_blocks * kTimelinePoints)) return 1;
  }
  return 0;
}

extern "C" int mb_last_k2_timeline(const mb_handle* h, double* us, int32_t cap_blocks) {
  if (!h || !us) { fail("mb_last_k2_timeline: null argument"); return -1; }
  if (!h->d_timeline || h->timeline_blocks <= 0) return 0;
  const int nb = std::min(h->timeline_blocks, (int)cap_blocks);
  std::vector<unsigned long long> t((size_t)h->timeline_blocks * kTimelinePoints);
  if (cudaMemcpy(t.data(), h->d_timeline, sizeof(unsigned long long) * t.size(), cudaMemcpyDeviceToHost) != cudaSuccess) {
    fail("mb_last_k2_timeline: copy failed");
    return -1;
  }
  unsigned long long t0 = ~0ull;
  for (int b = 0; b < h->timeline_blocks; ++b) t0 = std::min(t0, t[(size_t)b * kTimelinePoints]);
  for (int b = 0; b < nb; ++b)
    for (int k = 0; k < kTimelinePoints; ++k) {
      unsigned long long v = t[(size_t)b * kTimelinePoints + k];
      us[(size_t)b * kTimelinePoints + k] = v >= t0 && v != 0 ? (double)(v - t0) * 1e-3 : -1.0;
    }
  return nb;
}
extern "C" int mb_last_kernel_ms(const mb_handle* h, float ms[MB_NKERNELS]) {
  if (!h) return fail("null handle");
  memcpy(ms, h->last_ms, sizeof h->last_ms);
  return 0;
}

// K2 is launched with programmatic stream serialization: its blocks may be placed while the
// preceding kernel (K0 / K1) is still running and wait inside the kernel (griddepcontrol.wait)
template <int MODE, typename TF = double, int OB = 0>
static cudaError_t launch_k2(int NW, int grid, size_t smem, cudaStream_t st, const K2Args& a) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)k2_threads(NW));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  static const bool no_pdl = getenv("MB_NO_PDL") && atoi(getenv("MB_NO_PDL"));  // A/B switch (measurements)
  cfg.attrs = &attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  switch (NW) {
    case 1: return cudaLaunchKernelEx(&cfg, mb_k2_stars<1, MODE, TF, OB>, a);
    case 2: return cudaLaunchKernelEx(&cfg, mb_k2_stars<2, MODE, TF, OB>, a);
    default: return cudaLaunchKernelEx(&cfg, mb_k2_stars<4, MODE, TF, OB>, a);
  }
}

template <int MODE, typename TF = double, int OB = 0>
static int k2_occ(int NW, size_t smem) {
  int n = 0;
  switch (NW) {
    case 1: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, mb_k2_stars<1, MODE, TF, OB>, k2_threads(1), smem); break;
    case 2: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, mb_k2_stars<2, MODE, TF, OB>, k2_threads(2), smem); break;
    default: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, mb_k2_stars<4, MODE, TF, OB>, k2_threads(4), smem); break;
  }
  return n;
}
static int k2_blocks_per_sm(int kmode, int NW, size_t smem, int table_bytes, int ob) {
  if (kmode == kModeFactor && ob == 4) return k2_occ<kModeFactor, double, 4>(NW, smem);
  if (kmode == kModeFactor && ob == 6) return k2_occ<kModeFactor, double, 6>(NW, smem);
  if (kmode == kModeFactor && table_bytes == 4) return k2_occ<kModeFactor, float>(NW, smem);
  return kmode == kModeFactor ? k2_occ<kModeFactor>(NW, smem)
         : kmode == kModeTable ? k2_occ<kModeTable>(NW, smem) : k2_occ<kModeTableGrid>(NW, smem);
}

// h_phi (optional): host copy of phi, carried in K2's kernel parameters when the tables are built
// inside K2 and it is small enough; d_phi must then still be readable by the device (mapped memory)
static int eval_device(mb_handle* h, const double* d_phi, int W, int ndim, const double* const d_tab[MB_NPARAM],
                       double* d_out, cudaStream_t st, const double* h_phi = nullptr,
                       unsigned long long* d_flag = nullptr, unsigned long long flag_seq = 0) {
  if (W <= 0) return 0;
  if (ndim != h->ndim) return fail("phi has %d columns, the model has %d free variables", ndim, h->ndim);
  CK(cudaSetDevice(h->device));
  const int kmode = h->mode == MB_MODE_FACTOR ? kModeFactor : h->mode == MB_MODE_TABLE_UNIQUE ? kModeTable : kModeTableGrid;
  const bool elementwise = h->mode == MB_MODE_ELEMENTWISE;  // K1e (+ K1b) instead of K0 + K1; K2 gathers by grid row
  const bool prof = h->profiling != 0;
  float acc_ms[MB_NKERNELS] = {0, 0, 0};
  int launches = 0;
  const int nrows = h->nO + h->nI;
  for (int w0 = 0; w0 < W;) {
    int rem = W - w0;
    int NW = rem <= 32 ? 1 : rem <= 64 ? 2 : 4;
    while (NW > 1 && k2_smem_plan(kmode, h->nO, h->nI, h->nbins, 32 * NW, h->table_bytes, h->max_smem).total > (size_t)h->max_smem) NW >>= 1;
    const int Wpad = 32 * NW;
    const int wp = std::min(rem, Wpad);
    h->last_wp = wp;
    double* FO = h->d_F;                                // outer rows (TABLE modes: the age rows FA)
    double* FI = h->d_F + (size_t)h->nO * Wpad;         // inner rows (TABLE modes: the dust rows FB)
    double* Fage = h->d_F + (size_t)nrows * Wpad;       // plain age rows

    // FACTOR mode with fp64 tables and analytic models: K2 builds the factor tables in its own
    // prologue (look-up rows in the stage area), no K0 launch
    // host phi travels in the kernel parameters and is staged in shared memory; a batch too large
    // for that keeps K0 (one read of the mapped host memory per walker instead of one per block)
    const bool inline_phi = h_phi && (int64_t)wp * ndim <= kPhiInline;
    const size_t fuse_scratch = k2_fuse_scratch(h->nscratch, h->ncls_total, nrows, Wpad, (size_t)wp * ndim);
    // (measured: letting every block read a host phi from the mapped pinned buffer instead costs 24 us per
    // call -- 148 blocks x 4.6 KB over PCIe -- against the 11.7 KB parameter block's < 1 us)
    const bool may_fuse = kmode == kModeFactor && h->table_bytes == 8 && !d_tab && !h->no_fuse &&
                          h->ncls_total < 0x7fff && (!h_phi || inline_phi);
    // most (age, Z) bins a block can get: the launch has at least min(nbins, SMs) blocks
    const int bins_per_block = h->nbins > 0 ? (h->nbins + std::min(h->nbins, h->sm_count) - 1) / std::min(h->nbins, h->sm_count) : 0;
    const K2Smem plan = k2_smem_plan(kmode, h->nO, h->nI, h->nbins, Wpad, h->table_bytes, h->max_smem,
                                     may_fuse ? fuse_scratch : 0, h->nscratch, h->nOd * h->nI, bins_per_block, elementwise);
    const bool fused = may_fuse && plan.fuse_fits;
    if (prof) CK(cudaEventRecord(h->ev[0], st));
    if (!elementwise && !fused) {
      K0Args k0{};
      k0.model = h->mdev; k0.clsx = h->d_clsx; k0.phi = d_phi + (size_t)w0 * ndim;
      for (int p = 0; p < MB_NPARAM; ++p)
        k0.tab[p] = (d_tab && d_tab[p]) ? d_tab[p] + (size_t)w0 * h->mdev.ncls[p] : nullptr;
      k0.F = h->d_F; k0.rowsrc = h->d_rowsrc; k0.nrows = nrows; k0.nAge = h->nA; k0.W = wp; k0.Wpad = Wpad; k0.ndim = ndim;
      k0.ncls_total = h->ncls_total;
      mb_k0_factors<<<Wpad, kK0Threads, sizeof(double) * std::max(1, h->ncls_total), st>>>(k0);
      ++launches;
    }
    if (prof) CK(cudaEventRecord(h->ev[1], st));
    if (kmode != kModeFactor) {
      int64_t need = h->U * Wpad;
      if (need > h->cap_T) {
        CK(cudaStreamSynchronize(st));
        if (h->d_T) { CK(cudaFree(h->d_T)); h->device_bytes -= h->cap_T * 8; }
        h->d_T = nullptr; h->cap_T = 0;
        if (dev_alloc(h, &h->d_T, need)) return 1;
        h->cap_T = need;
      }
    }
    if (elementwise) {
      K1eArgs k1{};
      k1.model = h->mdev;
      for (int p = 0; p < MB_NPARAM; ++p) {
        k1.col[p] = h->d_col[p];
        k1.tab[p] = (d_tab && d_tab[p]) ? d_tab[p] + (size_t)w0 * (size_t)h->G : nullptr;
      }
      k1.phi = d_phi + (size_t)w0 * ndim; k1.T = h->d_T; k1.G = h->G;
      k1.W = wp; k1.Wpad = Wpad; k1.ndim = ndim;
      if (h->poisson) {
        k1.rowlist = h->d_binrows; k1.tile_start = h->d_tile_start; k1.gw = h->d_gw; k1.compl_ = h->d_compl;
        k1.partial = h->d_partial; k1.Fage = h->d_Fage; k1.bin_age = h->d_bin_age; k1.nbins = h->nbins;
        k1.ntiles = h->ntiles;
      } else {
        k1.ntiles = (h->G + kK1eRows - 1) / kK1eRows;
      }
      const size_t sm1 = k1e_smem_bytes(Wpad, ndim);
      const int blocks1 = (int)std::min<int64_t>(k1.ntiles, (int64_t)h->sm_count * 6);
      bool ln = true;  // every free dust parameter a single lognormal: the register-resident fast path
      for (int p = MB_P_AV; p < MB_NPARAM; ++p)
        ln = ln && (h->mdev.kind[p] == MB_FIXED || h->mdev.kind[p] == MB_LOGNORMAL);
      if (ln) mb_k1_elementwise<true><<<blocks1, kK1eThreads, sm1, st>>>(k1);
      else mb_k1_elementwise<false><<<blocks1, kK1eThreads, sm1, st>>>(k1);
      ++launches;
      if (h->poisson) {
        mb_k1b_binsums<<<h->nbins, kK1bThreads, 0, st>>>(h->d_partial, h->d_bin_tile, h->d_SC, Wpad);
        ++launches;
      }
    } else if (kmode != kModeFactor) {
      int64_t total = h->U * (Wpad / 2);
      int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)h->sm_count * 16);
      mb_k1_expand<<<blocks, 256, 0, st>>>(FO, FI, h->mode == MB_MODE_TABLE_GRID ? h->d_rowcode : nullptr,
                                           h->d_T, h->U, h->nB, Wpad);
      ++launches;
    }
    if (prof) CK(cudaEventRecord(h->ev[2], st));
    size_t smem = plan.total;
    // occupancy query cached per pass shape (it costs microseconds of host time per call)
    int& cached = h->occ_cache[NW == 1 ? 0 : NW == 2 ? 1 : 2];
    if (cached <= 0 || h->occ_smem[NW == 1 ? 0 : NW == 2 ? 1 : 2] != smem) {
      cached = k2_blocks_per_sm(kmode, NW, smem, h->table_bytes, h->ob);
      h->occ_smem[NW == 1 ? 0 : NW == 2 ? 1 : 2] = smem;
    }
    int per_sm = cached;
    if (per_sm < 1) return fail("K2 cannot be resident with %zu bytes of shared memory", smem);
    const int kK2Warps = k2_threads(NW) / 32;
    // at least one block even without entries (the result rows are written by K2's last block), and
    // enough blocks to spread the (age, Z) bins of the N-stars term
    int grid = std::min(h->sm_count * per_sm,
                        std::max(h->poisson ? std::min(h->nbins, h->sm_count) : 1, (h->nseg[NW == 4] + kK2Warps - 1) / kK2Warps));
    grid = std::max(1, std::min(grid, h->max_blocks));
    K2Args k2{};
    k2.ea = h->d_ea; k2.ec = h->c16 ? (const void*)h->d_ec16 : (const void*)h->d_ec; k2.seg = h->d_seg[NW == 4]; k2.nseg = h->nseg[NW == 4]; k2.FA = FO; k2.T = h->d_T;
    k2.gacc = h->d_gacc; k2.counters = h->d_counter; k2.out = d_out;
    k2.nA = h->nO; k2.nB = h->nI; k2.nOd = h->nOd; k2.Wpad = Wpad; k2.W = wp; k2.w0 = w0; k2.Wtot = W;
    k2.tab_off = (int32_t)plan.front;
    k2.scratch_off = (int32_t)plan.scratch_off; k2.separate = plan.separate ? 1 : 0;
    static const bool no_tma = getenv("MB_NO_TMA") && atoi(getenv("MB_NO_TMA"));  // A/B switch: per-lane cp.async ring
    k2.use_tma = no_tma ? 0 : 1;
    k2.slab_off = (int32_t)plan.slab_off;
    k2.pois.Ht = h->d_H; k2.pois.bin_agecls = h->d_bin_agecls; k2.pois.coef = h->d_coef; k2.pois.Pb = h->d_Pb;
    k2.pois.n_stars = (double)h->S_total; k2.pois.nbins = h->poisson ? h->nbins : 0;
    k2.pois.lgamma_n1 = std::lgamma((double)h->S_total + 1.0);
    k2.pois.nwarps = plan.pois_warps; k2.pois.scratch_off = (int32_t)plan.pois_off;
    k2.pois.warp_private = plan.warp_private ? 1 : 0; k2.pois.hts_off = (int32_t)plan.hts_off; k2.pois.hts_slices = plan.hts_slices;
    static const bool bins_last = getenv("MB_BINS_LAST") && atoi(getenv("MB_BINS_LAST"));  // A/B switch (measurements)
    k2.pois.bins_first = bins_last ? 0 : 1;
    k2.pois.SC = elementwise ? h->d_SC : nullptr;
    // plain age factors (the `ageprior` of helpers.py:135): K1e's per-bin rows, K0's rows in global
    // memory, or -- tables built inside K2 -- the outer table itself / the scratch look-up rows
    if (elementwise) k2.pois.fage = h->d_Fage;
    else if (!fused) k2.pois.fage = Fage;
    else {
      k2.pois.fage = nullptr;
      k2.pois.fage_smem_off = h->age_slot < 0 ? (int32_t)plan.front
                                               : (int32_t)(plan.scratch_off + (size_t)h->age_slot * Wpad * 8);
    }
    if (d_flag && w0 + wp >= W) { k2.done_flag = d_flag; k2.done_seq = flag_seq; }  // the last pass publishes
    k2.build_tables = fused ? 1 : 0;
    if (fused) {
      k2.model = h->mdev; k2.clsx = h->d_clsx; k2.ndim = ndim; k2.ncls_total = h->ncls_total;
      k2.rowsrc = h->d_rowslot; k2.lutslot = h->d_lutslot; k2.nscratch = h->nscratch;
      k2.phi = d_phi + (size_t)w0 * ndim;
      if (inline_phi) {
        k2.phi_inline = 1;
        memcpy(k2.phi_in, h_phi + (size_t)w0 * ndim, sizeof(double) * (size_t)wp * ndim);
      }
    }
    k2.timeline = nullptr;
    if (prof && h->d_timeline) {
      CK(cudaMemsetAsync(h->d_timeline, 0, sizeof(unsigned long long) * (size_t)grid * kTimelinePoints, st));
      k2.timeline = h->d_timeline;
      h->timeline_blocks = grid;
    }
    k2.peer.world = h->peer_world; k2.peer.rank = h->peer_rank; k2.peer.timeout_ns = h->peer_timeout_ns;
    if (h->peer_world > 1) {
      k2.peer.seq = ++h->peer_seq;
      for (int p = 0; p < h->peer_world; ++p) k2.peer.xbuf[p] = h->peer_xbuf[p];
    }
    if (kmode == kModeFactor && h->ob == 4) CK((launch_k2<kModeFactor, double, 4>(NW, grid, smem, st, k2)));
    else if (kmode == kModeFactor && h->ob == 6) CK((launch_k2<kModeFactor, double, 6>(NW, grid, smem, st, k2)));
    else if (kmode == kModeFactor && h->table_bytes == 4) CK((launch_k2<kModeFactor, float>(NW, grid, smem, st, k2)));
    else if (kmode == kModeFactor) CK(launch_k2<kModeFactor>(NW, grid, smem, st, k2));
    else if (kmode == kModeTable) CK(launch_k2<kModeTable>(NW, grid, smem, st, k2));
    else CK(launch_k2<kModeTableGrid>(NW, grid, smem, st, k2));
    ++launches;
    if (prof) {
      CK(cudaEventRecord(h->ev[3], st));
      CK(cudaEventSynchronize(h->ev[3]));
      for (int i = 0; i < MB_NKERNELS; ++i) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]));
        acc_ms[i] += ms;
      }
    }
    w0 += wp;
  }
  CK(cudaGetLastError());
  h->last_launches = launches;
  if (prof) memcpy(h->last_ms, acc_ms, sizeof acc_ms);
  return 0;
}

extern "C" int mb_lnlike_device(mb_handle* h, const double* d_phi, int32_t W, int32_t ndim, double* d_out,
                                void* stream) {
  if (!h || !d_phi || !d_out) return fail("mb_lnlike_device: null argument");
  for (int p = 0; p < MB_NPARAM; ++p)
    if (h->model.kind[p] == MB_TABULATED) return fail("model has tabulated parameters: use mb_lnlike_tabulated");
  return eval_device(h, d_phi, W, ndim, nullptr, d_out, (cudaStream_t)stream);
}

static int ensure(mb_handle* h, double** p, int64_t* cap, int64_t need) {
  if (need <= *cap) return 0;
  if (*p) { CK(cudaFree(*p)); h->device_bytes -= *cap * 8; }
  *p = nullptr; *cap = 0;
  if (dev_alloc(h, p, need)) return 1;
  *cap = need;
  return 0;
}

extern "C" int mb_lnlike_begin(mb_handle* h, const double* phi, int32_t W, int32_t ndim,
                               const double* const tab[MB_NPARAM]) {
  if (!h) return fail("mb_lnlike_begin: null handle");
  if (h->pending_W) return fail("mb_lnlike_begin: previous evaluation not collected with mb_lnlike_end");
  if (W <= 0) return fail("mb_lnlike_begin: W must be positive");
  if (ndim != h->ndim) return fail("phi has %d columns, the model has %d free variables", ndim, h->ndim);
  if (ndim > 0 && !phi) return fail("mb_lnlike: phi is NULL");
  if (!h->watches.empty() && check_watches(h)) return 1;
  CK(cudaSetDevice(h->device));
  const int64_t nphi = (int64_t)W * std::max(1, ndim), nout = (int64_t)MB_OUT_ROWS * W;
  // phi (a few KB) and the result rows (3 W doubles) travel through MAPPED pinned host memory:
  // K0 reads phi over PCIe, K2 writes the rows back, no separate copy operations in the stream
  if (nphi + nout > h->cap_pin) {
    CK(cudaStreamSynchronize(h->stream));
    if (h->h_pin) CK(cudaFreeHost(h->h_pin));
    h->h_pin = nullptr; h->cap_pin = 0;
    // (+ one 8-byte slot at the end: the completion flag)
    CK(cudaHostAlloc((void**)&h->h_pin, sizeof(double) * (size_t)(nphi + nout + 1), cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void**)&h->d_pin, h->h_pin, 0));
    h->cap_pin = nphi + nout;
    h->h_flag = reinterpret_cast<volatile unsigned long long*>(h->h_pin + h->cap_pin);
    h->d_flag = reinterpret_cast<unsigned long long*>(h->d_pin + h->cap_pin);
    *h->h_flag = 0ull;
  }
  double* hp = h->h_pin;
  double* ho = h->h_pin + nphi;
  if (ndim > 0) memcpy(hp, phi, sizeof(double) * (size_t)W * ndim);
  const double* dtab[MB_NPARAM] = {nullptr, nullptr, nullptr, nullptr};
  for (int p = 0; p < MB_NPARAM; ++p) {
    if (h->model.kind[p] != MB_TABULATED) continue;
    if (!tab || !tab[p]) return fail("parameter %d is tabulated but no table was passed", p);
    int64_t n = (int64_t)W * h->mdev.ncls[p];
    if (ensure(h, &h->d_tab[p], &h->cap_tab[p], n)) return 1;
    CK(cudaMemcpyAsync(h->d_tab[p], tab[p], sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    dtab[p] = h->d_tab[p];
  }
  bool any_tab = false;
  for (int p = 0; p < MB_NPARAM; ++p) any_tab |= dtab[p] != nullptr;
  const bool flag = h->use_flag && !h->profiling && h->h_flag;
  const unsigned long long seq = flag ? ++h->flag_seq : 0ull;
  if (eval_device(h, h->d_pin, W, ndim, any_tab ? dtab : nullptr, h->d_pin + nphi, h->stream, hp,
                  flag ? h->d_flag : nullptr, seq)) return 1;
  if (!flag) h->flag_seq = 0;  // mb_lnlike_end synchronizes the stream
  h->pending_W = W;
  h->pending_out = ho;
  return 0;
}

extern "C" int mb_lnlike_end(mb_handle* h, double* rows) {
  if (!h || !rows) return fail("mb_lnlike_end: null argument");
  if (!h->pending_W) return fail("mb_lnlike_end: no evaluation in flight");
  CK(cudaSetDevice(h->device));
  const int W = h->pending_W;
  h->pending_W = 0;
  bool seen = false;
  if (h->flag_seq && h->h_flag) {
    // poll the flag the kernel raises after its rows; look at the stream now and then so that a
    // failed launch or a fault cannot leave us spinning
    const unsigned long long want = h->flag_seq;
    for (unsigned spins = 0;; ++spins) {
      if (*h->h_flag == want) { seen = true; break; }
      if ((spins & 0x3fffu) == 0x3fffu) {
        cudaError_t q = cudaStreamQuery(h->stream);
        if (q != cudaErrorNotReady) {  // finished (flag will be there) or failed
          if (q != cudaSuccess) return fail("likelihood kernels failed: %s", cudaGetErrorString(q));
          break;
        }
      }
    }
  }
  if (!seen) CK(cudaStreamSynchronize(h->stream));
  memcpy(rows, h->pending_out, sizeof(double) * (size_t)MB_OUT_ROWS * W);
  return 0;
}

extern "C" void mb_combine(const double* rows, int32_t W, double* lnlike, int32_t* status) {
  for (int w = 0; w < W; ++w) {
    double tot = rows[MB_OUT_POISSON * W + w] + (rows[MB_OUT_STARSUM * W + w] + rows[MB_OUT_STARSUM_LO * W + w]);
    int st = MB_STATUS_OK;
    if (rows[MB_OUT_NONFINITE * W + w] > 0.0) st |= MB_STATUS_NONFINITE_STAR;
    if (rows[MB_OUT_NONFINITE * W + w] < 0.0) st |= MB_STATUS_PEER_TIMEOUT;
    if (tot == 0.0) st |= MB_STATUS_TOTAL_ZERO;
    lnlike[w] = tot;
    status[w] = st;
  }
}

extern "C" int mb_lnlike_tabulated(mb_handle* h, const double* phi, int32_t W, int32_t ndim,
                                   const double* const tab[MB_NPARAM], double* lnlike, int32_t* status) {
  if (!h || !lnlike || !status) return fail("mb_lnlike: null argument");
  if (W <= 0) return 0;
  if (mb_lnlike_begin(h, phi, W, ndim, tab)) return 1;
  std::vector<double> rows((size_t)MB_OUT_ROWS * W);
  if (mb_lnlike_end(h, rows.data())) return 1;
  mb_combine(rows.data(), W, lnlike, status);
  return 0;
}

extern "C" int mb_lnlike(mb_handle* h, const double* phi, int32_t W, int32_t ndim, double* lnlike,
                         int32_t* status) {
  return mb_lnlike_tabulated(h, phi, W, ndim, nullptr, lnlike, status);
}

// lnprob of the reference (singlepop_dust_model.py:277-304) for a batch with flat (box) priors:
// 0 inside [lo, hi] and -inf outside (:255-271; a NaN compares false on both sides and passes, as
// upstream); the likelihood is evaluated only for the walkers inside the box (:302-303).
extern "C" int mb_lnprob_box(mb_handle* h, const double* phi, int32_t W, int32_t ndim, const double* lo,
                             const double* hi, double* lnprob, int32_t* status, int32_t* last_inside) {
  if (!h || !lnprob || !status) return fail("mb_lnprob_box: null argument");
  if (W <= 0) return 0;
  if (ndim != h->ndim) return fail("phi has %d columns, the model has %d free variables", ndim, h->ndim);
  if (ndim > 0 && (!phi || !lo || !hi)) return fail("mb_lnprob_box: null array");
  std::vector<int32_t>& idx = h->scratch_idx;
  idx.clear();
  for (int w = 0; w < W; ++w) {
    const double* p = phi + (size_t)w * ndim;
    bool inside = true;
    for (int k = 0; k < ndim; ++k)
      if (p[k] < lo[k] || p[k] > hi[k]) { inside = false; break; }
    if (inside) idx.push_back(w);
    lnprob[w] = -INFINITY;
    status[w] = MB_STATUS_OK;
  }
  if (last_inside) *last_inside = idx.empty() ? -1 : idx.back();
  const int Win = (int)idx.size();
  if (Win == 0) return 0;
  const double* src = phi;
  if (Win != W) {  // compact the walkers that are evaluated
    h->scratch_phi.resize((size_t)Win * ndim);
    for (int i = 0; i < Win; ++i)
      memcpy(&h->scratch_phi[(size_t)i * ndim], phi + (size_t)idx[(size_t)i] * ndim, sizeof(double) * ndim);
    src = h->scratch_phi.data();
  }
  if (mb_lnlike_begin(h, src, Win, ndim, nullptr)) return 1;
  h->scratch_rows.resize((size_t)MB_OUT_ROWS * Win);
  if (mb_lnlike_end(h, h->scratch_rows.data())) return 1;
  h->scratch_val.resize((size_t)Win);
  h->scratch_st.resize((size_t)Win);
  mb_combine(h->scratch_rows.data(), Win, h->scratch_val.data(), h->scratch_st.data());
  for (int i = 0; i < Win; ++i) {
    lnprob[idx[(size_t)i]] = 0.0 + h->scratch_val[(size_t)i];
    status[idx[(size_t)i]] = h->scratch_st[(size_t)i];
  }
  return 0;
}

extern "C" int mb_lnlike_pixels(mb_handle* h, const double* phi, int32_t W, int32_t ndim, double* lnlike,
                                int32_t* status) {
  return mb_lnlike_pixels_tabulated(h, phi, W, ndim, nullptr, lnlike, status);
}

// Pixel batch from host arrays.  lo/hi/park (all or none): the flat prior box of mb_lnprob_pixels_box -- a
// walker outside [lo, hi] is evaluated at `park` (it occupies a column either way) and reported as -inf.
static int pixels_host_eval(mb_handle* h, const double* phi, int32_t W, int32_t ndim, const double* const tab[MB_NPARAM],
                            double* lnlike, int32_t* status, const double* lo, const double* hi, const double* park);

extern "C" int mb_lnlike_pixels_tabulated(mb_handle* h, const double* phi, int32_t W, int32_t ndim,
                                          const double* const tab[MB_NPARAM], double* lnlike, int32_t* status) {
  return pixels_host_eval(h, phi, W, ndim, tab, lnlike, status, nullptr, nullptr, nullptr);
}

extern "C" int mb_lnprob_pixels_box(mb_handle* h, const double* phi, int32_t W, int32_t ndim, const double* lo,
                                    const double* hi, const double* park, double* lnprob, int32_t* status) {
  if (ndim > 0 && (!lo || !hi || !park)) return fail("mb_lnprob_pixels_box: null box array");
  for (int k = 0; k < ndim; ++k)
    if (!(park[k] >= lo[k] && park[k] <= hi[k])) return fail("mb_lnprob_pixels_box: park[%d] is outside the box", k);
  return pixels_host_eval(h, phi, W, ndim, nullptr, lnprob, status, lo, hi, park);
}

static int pixels_host_eval(mb_handle* h, const double* phi, int32_t W, int32_t ndim, const double* const tab[MB_NPARAM],
                            double* lnlike, int32_t* status, const double* lo, const double* hi, const double* park) {
  if (!h || !lnlike || !status) return fail("mb_lnlike_pixels: null argument");
  if (h->P <= 0) return fail("handle was not created with mb_create_pixels");
  if (ndim != h->ndim) return fail("phi has %d columns, the model has %d free variables", ndim, h->ndim);
  if (ndim > 0 && !phi) return fail("mb_lnlike_pixels: phi is NULL");
  if (W < 1) return fail("mb_lnlike_pixels: W must be positive");
  if (!h->watches.empty() && check_watches(h)) return 1;
  CK(cudaSetDevice(h->device));
  const int64_t ncol = h->P * W;
  const int64_t nbox = lo ? 3 * (int64_t)ndim : 0;  // lo, hi, park travel behind phi in the same copy
  const int64_t nphi = ncol * std::max(1, ndim) + nbox, nout = (int64_t)MB_OUT_ROWS * ncol;
  if (ensure(h, &h->d_phi, &h->cap_phi, nphi)) return 1;
  if (ensure(h, &h->d_out, &h->cap_out, nout)) return 1;
  if (nphi + nout > h->cap_pin) {
    if (h->h_pin) CK(cudaFreeHost(h->h_pin));
    h->h_pin = nullptr; h->cap_pin = 0;
    CK(cudaHostAlloc((void**)&h->h_pin, sizeof(double) * (size_t)(nphi + nout + 1), cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void**)&h->d_pin, h->h_pin, 0));
    h->cap_pin = nphi + nout;
    h->h_flag = reinterpret_cast<volatile unsigned long long*>(h->h_pin + h->cap_pin);  // (slot kept for mb_lnlike_begin)
    h->d_flag = reinterpret_cast<unsigned long long*>(h->d_pin + h->cap_pin);
    *h->h_flag = 0ull;
  }
  double* hp = h->h_pin;
  double* ho = h->h_pin + nphi;
  if (ndim > 0) memcpy(hp, phi, sizeof(double) * (size_t)ncol * ndim);
  const double* d_box = nullptr;
  if (nbox > 0) {
    double* hb = hp + nphi - nbox;
    memcpy(hb, lo, sizeof(double) * ndim);
    memcpy(hb + ndim, hi, sizeof(double) * ndim);
    memcpy(hb + 2 * ndim, park, sizeof(double) * ndim);
    d_box = h->d_phi + (nphi - nbox);
  }
  CK(cudaMemcpyAsync(h->d_phi, hp, sizeof(double) * (size_t)nphi, cudaMemcpyHostToDevice, h->stream));
  // host-evaluated tables of MB_TABULATED parameters: [P * W][n_classes]
  const double* dtab[MB_NPARAM] = {nullptr, nullptr, nullptr, nullptr};
  bool any_tab = false;
  for (int p = 0; p < MB_NPARAM; ++p) {
    if (h->model.kind[p] != MB_TABULATED) continue;
    if (!tab || !tab[p]) return fail("parameter %d is tabulated but no table was passed", p);
    const int64_t n = ncol * h->mdev.ncls[p];
    if (ensure(h, &h->d_tab[p], &h->cap_tab[p], n)) return 1;
    CK(cudaMemcpyAsync(h->d_tab[p], tab[p], sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    dtab[p] = h->d_tab[p];
    any_tab = true;
  }
  if (eval_pixels_device(h, h->d_phi, W, ndim, h->d_out, h->stream, any_tab ? dtab : nullptr, d_box)) return 1;
  CK(cudaMemcpyAsync(ho, h->d_out, sizeof(double) * (size_t)nout, cudaMemcpyDeviceToHost, h->stream));
  // While the device works: which columns are outside the box?  K0 makes the same IEEE compares on the same
  // values and builds those columns from `park`; here they are only listed, to be reported as -inf.
  std::vector<int32_t>& outside = h->scratch_idx;
  outside.clear();
  if (nbox > 0) {
    for (int64_t c = 0; c < ncol; ++c) {
      const double* p = phi + (size_t)c * ndim;
      bool inside = true;
      for (int k = 0; k < ndim; ++k) inside &= !(p[k] < lo[k] || p[k] > hi[k]);  // same compare as mb_lnprob_box
      if (!inside) outside.push_back((int32_t)c);
    }
  }
  CK(cudaStreamSynchronize(h->stream));
  for (int64_t p = 0; p < h->P; ++p)  // rows are [p][3][W]
    mb_combine(ho + (size_t)p * MB_OUT_ROWS * W, W, lnlike + (size_t)p * W, status + (size_t)p * W);
  for (int32_t c : outside) {  // singlepop_dust_model.py:302-303: never evaluated in the reference
    lnlike[c] = -INFINITY;
    status[c] = MB_STATUS_OK;
  }
  return 0;
}

extern "C" int mb_likelihoods_from_lnp(int32_t device, int64_t K, int64_t S, const int64_t* indxs, double* vals,
                                       int64_t G, const double* pw) {
  if (K < 0 || S < 0 || G <= 0 || !pw) return fail("mb_likelihoods_from_lnp: bad argument");
  int64_t n = K * S;
  if (n == 0) return 0;
  if (!indxs || !vals) return fail("mb_likelihoods_from_lnp: null array");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("no CUDA device available; megabeast_b200 has no CPU fallback");
  CK(cudaSetDevice(device));
  int64_t* d_i = nullptr; double* d_v = nullptr; double* d_p = nullptr; int* d_e = nullptr;
  int rc = 0, herr = 0;
  auto body = [&]() -> int {
    CK(cudaMalloc((void**)&d_i, sizeof(int64_t) * (size_t)n));
    CK(cudaMalloc((void**)&d_v, sizeof(double) * (size_t)n));
    CK(cudaMalloc((void**)&d_p, sizeof(double) * (size_t)G));
    CK(cudaMalloc((void**)&d_e, sizeof(int)));
    CK(cudaMemset(d_e, 0, sizeof(int)));
    CK(cudaMemcpy(d_i, indxs, sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_v, vals, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_p, pw, sizeof(double) * (size_t)G, cudaMemcpyHostToDevice));
    int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
    mb_ingest_lnp<<<blocks, 256>>>(n, d_i, d_v, G, d_p, d_e);
    CK(cudaGetLastError());
    CK(cudaMemcpy(vals, d_v, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&herr, d_e, sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
  };
  rc = body();
  cudaFree(d_i); cudaFree(d_v); cudaFree(d_p); cudaFree(d_e);
  if (rc) return rc;
  if (herr) return fail("index out of bounds for the prior_weight column");
  return 0;
}
