// engine.cu — the extern "C" boundary (include/sgdnet_b200.h) and the lambda-path driver behind it.
//
// What replaces what: sgdnet_fit_dense / sgdnet_fit_sparse are SgdnetDense / SgdnetSparse (reference
// src/sgdnet.cpp:359-375); sgdnet_fit_batch_* is the cv_sgdnet double loop (R/cv_sgdnet.R:160-200). The driver below
// is SetupSgdnet's lambda loop (src/sgdnet.cpp:217-273) turned into a per-fit state machine that lives on the device
// (Progress). Every fit of a batch is its own asynchronous pipeline on its own pair of streams:
//     [sampling indices (rng.cu) -> conflict codes] -> lag-scaling table (new lambda only) -> SAGA epochs
//     -> [debug epoch loss] -> deviance + rescale + archive (finish-lambda)
// The kernel that ends a step publishes the fit's Progress to pinned host memory; one host thread polls the fits'
// round ids and submits each fit's next launch as soon as its last one has published (no lock-step rounds, no stream
// synchronisation, any mix of kernel variants in one batch). The next launch's indices and conflict codes are
// prepared on the fit's second stream while the current launch solves. Warm-start state never leaves HBM
// (src/sgdnet.cpp:186-198). The design-level setup (CSC -> CSR, column statistics, scaling, row subsets, largest row
// norm, X^T y) runs on the device (setup.cu); the O(n) response statistics and the lambda grid stay on the host
// (host_setup.cu). With R's default generator the sampling sequence is produced on the device and the caller's
// generator is handed back advanced by exactly n * npasses draws; other generators (callback, fixed sequence) are
// drawn on the host in the reference's order, one launch ahead.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sgdnet_b200.h"
#include "host_setup.h"
#include <nvtx3/nvToolsExt.h>

#include "kernels.h"
#include "setup.h"

namespace sgd {

thread_local std::string g_error;

// Every fit of a batch runs on its own stream. The driver maps streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues
// (8 by default, 32 at most); streams that share a queue can hold each other up. Ask for the maximum unless the
// caller chose a value (only effective when the CUDA context is created after this library is loaded; harnesses that
// initialise CUDA first set the variable themselves).
struct ConnectionsEnv {
  ConnectionsEnv() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }
} g_connections_env;

struct CudaFail {
  cudaError_t e;
  const char* what;
};
#define CK(call)                                   \
  do {                                             \
    cudaError_t e_ = (call);                       \
    if (e_ != cudaSuccess) throw CudaFail{e_, #call}; \
  } while (0)

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// SGDNET_TIMING=1: host setup phases to stderr (development aid)
struct PhaseTimer {
  bool on = std::getenv("SGDNET_TIMING") != nullptr;
  double t = now_s();
  void lap(const char* what) {
    if (!on) return;
    const double u = now_s();
    std::fprintf(stderr, "[sgdnet_b200] %-28s %8.1f ms\n", what, (u - t) * 1e3);
    t = u;
  }
};

// Device and pinned host memory owned by one engine; freed together. Sub-allocated from a few large slabs: a batch of
// cv fits makes thousands of small allocations, and thousands of cudaMalloc / cudaFree / cudaMallocHost calls cost
// seconds (cudaFree synchronises the device every time).
struct Arena {
  struct Slab {
    char* base = nullptr;
    size_t size = 0, used = 0;
  };
  std::vector<Slab> dev, pin;
  cudaStream_t stream = nullptr;             // zero-fills are ordered on the engine's stream (it is non-blocking: the
                                             // legacy default stream would not be ordered with it)
  Arena() = default;
  Arena(const Arena&) = delete;              // owns raw device / pinned pointers
  Arena& operator=(const Arena&) = delete;
  static constexpr size_t kAlign = 256;
  static constexpr size_t kDevSlab = size_t(128) << 20, kPinSlab = size_t(8) << 20;

  void* carve(std::vector<Slab>& slabs, size_t bytes, size_t slab_size, bool pinned) {
    bytes = (std::max<size_t>(bytes, 1) + kAlign - 1) & ~(kAlign - 1);
    for (Slab& s : slabs)
      if (s.size - s.used >= bytes) {
        void* p = s.base + s.used;
        s.used += bytes;
        return p;
      }
    Slab s;
    s.size = std::max(bytes, slab_size);
    void* p = nullptr;
    if (pinned) CK(cudaMallocHost(&p, s.size));
    else CK(cudaMalloc(&p, s.size));
    s.base = static_cast<char*>(p);
    s.used = bytes;
    slabs.push_back(s);
    return p;
  }
  // one slab for everything a batch is about to allocate (a cudaMalloc per fit costs milliseconds)
  void reserve(size_t bytes) {
    if (bytes == 0) return;
    Slab s;
    s.size = bytes;
    void* p = nullptr;
    CK(cudaMalloc(&p, s.size));
    s.base = static_cast<char*>(p);
    dev.push_back(s);
  }
  template <typename T>
  T* alloc(size_t count, bool zero = true) {
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    void* p = carve(dev, bytes, kDevSlab, false);
    if (zero) CK(cudaMemsetAsync(p, 0, bytes, stream));
    return static_cast<T*>(p);
  }
  template <typename T>
  T* upload(const std::vector<T>& v) {
    T* p = alloc<T>(v.size(), false);
    if (!v.empty()) CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return p;
  }
  template <typename T>
  T* upload_from(const T* src, size_t count) {
    T* p = alloc<T>(count, false);
    if (count) CK(cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice));
    return p;
  }
  template <typename T>
  T* host(size_t count) {
    return static_cast<T*>(carve(pin, std::max<size_t>(count, 1) * sizeof(T), kPinSlab, true));
  }
  ~Arena() {
    for (Slab& s : dev) cudaFree(s.base);
    for (Slab& s : pin) cudaFreeHost(s.base);
  }
};

struct DeviceDesign {
  const double* xd = nullptr;
  const RowInfo* rows = nullptr;
  const int32_t* ci = nullptr;
  const double* cv = nullptr;
  const double *c = nullptr, *x_center = nullptr, *x_scale = nullptr;
  const int32_t* local = nullptr;     // sparse row subset: source row -> position in this view, or -1 (X^T y walks the CSC)
};

// The caller's matrix on the device: the dgCMatrix as given (CSC) plus its padded CSR (setup.cu), or the numeric
// matrix as given (column-major).
struct RawInput {
  bool sparse = false;
  int64_t n = 0, nnz = 0;
  int32_t p = 0;
  const int32_t *csc_i = nullptr, *csc_p = nullptr;
  const double* csc_x = nullptr;
  const RowInfo* rows = nullptr;
  const int32_t* ci = nullptr;
  const double* cv = nullptr;
  int64_t n_entries = 0;
  int64_t max_row_nnz = 0;
  const double* dense_cm = nullptr;
};

enum class Variant { Dense, SparseK1, SparseCentred, SparseGeneric };
enum class Phase { Idle, Solver, Finish, Parked, Done };

// One fit of a batch = one asynchronous pipeline: its own stream, its own rounds
//     [sampling indices -> conflict codes] -> lag-scaling table (new lambda) -> SAGA epochs
//     and, when a lambda is finished (or after every epoch in debug mode), [epoch loss] -> deviance + rescale + archive
// submitted as soon as the fit's previous launch has published its Progress to pinned host memory. Nothing makes a
// small fit wait for a large one.
struct FitJob {
  std::shared_ptr<HostDesign> design;
  DeviceDesign ddev;
  FitPlan plan;
  sgdnet_rng* rng = nullptr;
  FitDev dev{};                       // host mirror of the device struct
  FitDev* dev_ptr = nullptr;
  Progress* prog_ptr = nullptr;
  Progress* mirror = nullptr;         // pinned: written by the kernels (publish_progress), polled by the host
  Variant variant = Variant::Dense;
  // ---- index stream, host generators (callback / explicit sequence / MT when SGDNET_HOST_RNG is set)
  std::vector<uint32_t> pending;      // generated, not yet consumed
  size_t pending_head = 0;
  uint64_t consumed = 0;              // draws consumed by finished epochs
  std::vector<std::pair<uint64_t, sgdnet_rng>> marks;   // (draws generated before, generator state) per block
  uint64_t generated = 0;
  uint32_t* seq_pin = nullptr;
  // ---- index stream, device generator (R's Mersenne-Twister, rng.cu)
  bool device_rng = false;
  MtState* rng_dev = nullptr;         // the caller's generator as uploaded at the start of run()
  MtState* rng_pin = nullptr;
  const MtState* cur_state = nullptr; // generator state the next launch starts from (device)
  sgdnet_rng handed_back{};           // the generator as last returned to the caller (its prepared launch stays valid)
  bool have_handed_back = false;
  // ---- launch buffers, double buffered: [buf] is consumed by the launch in flight while [buf ^ 1] is prepared
  uint32_t* seq_dev[2] = {nullptr, nullptr};
  uint64_t* dep_dev[2] = {nullptr, nullptr};   // sparse K == 1: conflict codes of the staged sequence (wave_deps_kernel)
  uint8_t* dup_dev[2] = {nullptr, nullptr};
  MtState* snap_dev[2] = {nullptr, nullptr};   // [epochs_per_launch + 1] generator snapshots at epoch boundaries
  int epochs_per_launch = 1;
  int buf = 0;
  bool prepped = false;               // [buf] already holds the indices (+ codes) that start at cur_state
  bool prepped_next = false;          // [buf ^ 1] is being prepared on st_prep assuming this launch uses all its epochs
  // ---- pipeline
  cudaStream_t st = nullptr, st_prep = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_idx = nullptr, ev_prep = nullptr, ev_f0 = nullptr, ev_f1 = nullptr;
  Phase phase = Phase::Idle;
  uint32_t round_id = 0;
  int ne_submitted = 0;
  bool idx_on_prep = false;           // this launch's indices were produced on st_prep
  bool needs_finish = false;          // the last solver launch ended a lambda (or, debug mode, an epoch): passes are due
  bool stale_prep = false;            // a prepared launch was discarded; its kernels may still be running on st_prep
  nvtxRangeId_t nvtx_launch = 0, nvtx_lambda = 0;      // open NVTX ranges (one per solver launch, one per lambda)
  bool nvtx_lambda_open = false;
  uint16_t* pos_scratch = nullptr;    // sparse + virtual centring: position maps when the state does not fit shared memory
  int loss_blocks = 1;
  int loss_tiles = 0;                 // > 0: CTAs of the bulk-copy tile form of the loss pass (sparse K == 1)
  int mask_words = 0;                 // sparse K == 1: words of the nonzero-coefficient bitmap (0: p too large for it)
  size_t dense_smem = 0;
  bool dense_cluster = false;         // wide dense design: the cluster kernel (saga_dense_cluster.cu)
  double seconds_solver = 0.0, seconds_dev = 0.0;
  uint64_t launches = 0;
  double t_submit = 0.0, t_finish = 0.0;
  uint64_t solver_ns_seen = 0;
  // ---- scoring of held-out rows after the fit (cv)
  const int32_t* test_rows = nullptr;
  int64_t n_test = 0;
  int32_t measure = 0;
  double* score_dev = nullptr;
  bool scored = false;
  bool path_only = false;             // setup up to the lambda path only: no device state, nothing runs
};

struct Engine {
  Arena arena;
  RawInput raw;
  std::mutex xt_mutex;                // X^T y requests from the plan threads are served one at a time
  std::vector<double> y_cm;           // caller's y, n x Ky column-major
  int Ky = 1;
  std::map<std::pair<const int32_t*, int>, std::pair<std::shared_ptr<HostDesign>, DeviceDesign>> designs;
  std::vector<FitJob> jobs;
  cudaStream_t stream = nullptr;      // setup / scoring of single calls
  int sms = 148;
  bool trace_rounds = std::getenv("SGDNET_TRACE_ROUNDS") != nullptr;   // per-launch device times on stderr
  bool host_rng = std::getenv("SGDNET_HOST_RNG") != nullptr;           // draw MT indices on the host (development aid)
  bool no_overlap = std::getenv("SGDNET_NO_PREP_OVERLAP") != nullptr;  // prepare a launch only when it is due
  bool no_centred = std::getenv("SGDNET_NO_CENTRED") != nullptr;         // sparse + standardize on the generic kernel (measurement aid)
  bool no_loss_tiles = std::getenv("SGDNET_NO_LOSS_TILES") != nullptr;   // loss pass without the bulk-copy tile form (measurement aid)
  double seconds_setup = 0.0;
  double t_begin = 0.0, t_run = 0.0;
  size_t l2_persist_bytes = 0, l2_window_max = 0;   // L2 set aside for persisting lines (0: not available / switched off)
  // raw design for scoring (device)
  DeviceDesign raw_dev;
  bool raw_uploaded = false;
  double* yraw_dev = nullptr;
  std::map<const int32_t*, int32_t*> test_dev;

  int dev_id = 0;        // the device this engine lives on: helper threads must select it (a new host thread starts on device 0)
  Engine() {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) throw CudaFail{e == cudaSuccess ? cudaErrorNoDevice : e, "no CUDA device (no CPU fallback)"};
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    arena.stream = stream;
    cudaGetDevice(&dev_id);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
    // L2 persistence for the sparse coefficient records: measured on config 2, it changes neither the solver's time
    // (301.9 vs 300.7 ms per epoch) nor helps anything else, and the L2 set aside for it slows the streaming passes
    // (deviance pass 0.51 -> 0.96 ms). Off unless asked for.
    if (std::getenv("SGDNET_L2_PERSIST") != nullptr) {
      int max_persist = 0, max_window = 0;
      cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev_id);
      cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev_id);
      if (max_persist > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, static_cast<size_t>(max_persist)) == cudaSuccess) {
        l2_persist_bytes = static_cast<size_t>(max_persist);
        l2_window_max = static_cast<size_t>(max_window);
      }
      (void)cudaGetLastError();
    }
    t_begin = now_s();
  }
  ~Engine() {
    PhaseTimer pt;
    struct Lap {
      PhaseTimer& t;
      ~Lap() { t.lap("engine teardown"); }
    } lap{pt};
    for (FitJob& j : jobs) {
      for (cudaEvent_t ev : {j.ev0, j.ev1, j.ev_idx, j.ev_prep, j.ev_f0, j.ev_f1})
        if (ev) cudaEventDestroy(ev);
      if (j.st) cudaStreamDestroy(j.st);
      if (j.st_prep) cudaStreamDestroy(j.st_prep);
    }
    if (stream) cudaStreamDestroy(stream);
  }

  // ---------------------------------------------------------------------------------- input and designs (device)
  template <typename T>
  T download_scalar(const T* dev) {
    T v;
    CK(cudaMemcpyAsync(&v, dev, sizeof(T), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return v;
  }
  // pageable host memory -> device, ordered on `stream` (the DMA of a synchronous cudaMemcpy may still be in flight
  // when it returns, on the legacy stream; this one is followed by work on `stream`)
  template <typename T>
  T* upload_on_stream(const T* src, size_t count) {
    T* p = arena.alloc<T>(count, false);
    if (count) CK(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, stream));
    return p;
  }

  void load_sparse(const int32_t* ci_, const int32_t* cp_, const double* cx_, int64_t n_, int64_t p_) {
    PhaseTimer pt;
    raw.sparse = true;
    raw.n = n_;
    raw.p = static_cast<int32_t>(p_);
    raw.nnz = cp_[p_];
    raw.csc_i = upload_on_stream(ci_, static_cast<size_t>(raw.nnz));
    raw.csc_p = upload_on_stream(cp_, static_cast<size_t>(p_) + 1);
    raw.csc_x = upload_on_stream(cx_, static_cast<size_t>(raw.nnz));
    pt.lap("CSC upload (issued)");
    // CSC -> padded CSR on the device (AdaptiveTranspose, src/utils.h:276-281)
    int32_t* counts = arena.alloc<int32_t>(static_cast<size_t>(raw.n), false);
    int32_t* cursor = arena.alloc<int32_t>(static_cast<size_t>(raw.n), false);
    int64_t* block_tot = arena.alloc<int64_t>(static_cast<size_t>(csc_to_csr_scan_blocks(raw.n)), false);
    int64_t* totals = arena.alloc<int64_t>(2, false);
    CK(csc_to_csr_counts(raw.csc_i, raw.nnz, raw.n, counts, block_tot, totals, stream));
    int64_t tot[2];
    CK(cudaMemcpyAsync(tot, totals, sizeof(tot), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    raw.n_entries = tot[0] + 4;
    raw.max_row_nnz = tot[1];
    RowInfo* rows = arena.alloc<RowInfo>(static_cast<size_t>(raw.n), false);
    int32_t* ci = arena.alloc<int32_t>(static_cast<size_t>(raw.n_entries), false);
    double* cv = arena.alloc<double>(static_cast<size_t>(raw.n_entries), false);
    CK(cudaMemsetAsync(ci + tot[0], 0, 4 * sizeof(int32_t), stream));
    CK(cudaMemsetAsync(cv + tot[0], 0, 4 * sizeof(double), stream));
    int32_t* sk = nullptr;
    double* sv = nullptr;
    if (tot[1] > 128) {      // rows too long for the in-register sort use scratch of the same size
      sk = arena.alloc<int32_t>(static_cast<size_t>(raw.n_entries), false);
      sv = arena.alloc<double>(static_cast<size_t>(raw.n_entries), false);
    }
    CK(csc_to_csr_fill(raw.csc_i, raw.csc_p, raw.csc_x, raw.n, raw.p, counts, block_tot, rows, cursor, ci, cv, sk, sv, stream));
    raw.rows = rows;
    raw.ci = ci;
    raw.cv = cv;
    if (PhaseTimer().on) CK(cudaStreamSynchronize(stream));
    pt.lap("CSC -> CSR (device)");
  }

  void load_dense(const double* x, int64_t n_, int64_t p_) {
    PhaseTimer pt;
    raw.sparse = false;
    raw.n = n_;
    raw.p = static_cast<int32_t>(p_);
    raw.dense_cm = upload_on_stream(x, static_cast<size_t>(n_) * p_);
    pt.lap("dense upload (issued)");
  }

  std::pair<std::shared_ptr<HostDesign>, DeviceDesign> get_design(const int32_t* rows, int64_t n_rows, bool standardize) {
    auto key = std::make_pair(rows, standardize ? 1 : 0);
    auto it = designs.find(key);
    if (it != designs.end()) return it->second;
    auto hd = std::make_shared<HostDesign>();
    PhaseTimer pt;
    hd->sparse = raw.sparse;
    hd->standardized = standardize;
    hd->n = rows ? n_rows : raw.n;
    hd->p = raw.p;
    hd->ld = (raw.p + 1) & ~1;
    const int64_t n = hd->n;
    const int32_t p = raw.p;
    DeviceDesign dd;
    const int32_t* subset_dev = rows ? upload_on_stream(rows, static_cast<size_t>(n_rows)) : nullptr;
    double* x_center = arena.alloc<double>(p);                         // zero
    double* x_scale = arena.alloc<double>(p, false);
    double* c = arena.alloc<double>(p);                                // zero
    {
      std::vector<double> ones(p, 1.0);
      CK(cudaMemcpyAsync(x_scale, ones.data(), sizeof(double) * p, cudaMemcpyHostToDevice, stream));
      CK(cudaStreamSynchronize(stream));                              // `ones` goes away
    }
    unsigned long long* norm_bits = arena.alloc<unsigned long long>(1, false);
    if (raw.sparse) {
      int32_t* local = nullptr;
      if (rows) {
        local = arena.alloc<int32_t>(static_cast<size_t>(raw.n), false);
        CK(subset_local(subset_dev, n, raw.n, local, stream));
      }
      dd.local = local;
      if (standardize) CK(sparse_col_stats(raw.csc_i, raw.csc_p, raw.csc_x, p, local, n, x_center, x_scale, c, stream));
      if (!rows && !standardize) {
        dd.rows = raw.rows;       // all rows, values as given: the raw CSR is the design
        dd.ci = raw.ci;
        dd.cv = raw.cv;
      } else {
        int32_t* counts = arena.alloc<int32_t>(static_cast<size_t>(n), false);
        int32_t* cursor = arena.alloc<int32_t>(static_cast<size_t>(n), false);
        int64_t* block_tot = arena.alloc<int64_t>(static_cast<size_t>(csc_to_csr_scan_blocks(n)), false);
        int64_t* totals = arena.alloc<int64_t>(2, false);
        CK(subset_counts(raw.rows, subset_dev, n, counts, block_tot, totals, stream));
        const int64_t entries = download_scalar(totals);
        RowInfo* drows = arena.alloc<RowInfo>(static_cast<size_t>(n), false);
        int32_t* dci = arena.alloc<int32_t>(static_cast<size_t>(entries) + 4, false);
        double* dcv = arena.alloc<double>(static_cast<size_t>(entries) + 4, false);
        CK(cudaMemsetAsync(dci + entries, 0, 4 * sizeof(int32_t), stream));
        CK(cudaMemsetAsync(dcv + entries, 0, 4 * sizeof(double), stream));
        CK(gather_rows(raw.rows, raw.ci, raw.cv, subset_dev, n, counts, block_tot, drows, cursor, dci, dcv,
                       standardize ? x_scale : nullptr, stream));
        dd.rows = drows;
        dd.ci = dci;
        dd.cv = dcv;
      }
      CK(sparse_norm_max(dd.rows, dd.ci, dd.cv, n, p, standardize ? c : nullptr, norm_bits, stream));
    } else {
      double* xd = arena.alloc<double>(static_cast<size_t>(n) * hd->ld, false);
      CK(dense_design(raw.dense_cm, raw.n, p, subset_dev, n, hd->ld, standardize, x_center, x_scale, xd, norm_bits, stream));
      dd.xd = xd;
    }
    const unsigned long long bits = download_scalar(norm_bits);
    std::memcpy(&hd->norm_max, &bits, sizeof(double));
    dd.c = c;
    dd.x_center = x_center;
    dd.x_scale = x_scale;
    pt.lap("design (device)");
    // X^T y for LambdaMax: asked for by the plan (possibly from a helper thread), answered by the device
    const DeviceDesign ddc = dd;
    const int64_t n_view = n;
    const int32_t ld = hd->ld;
    const bool std_ = standardize;
    hd->xt_times_fn = [this, ddc, n_view, p, ld, std_](const std::vector<double>& ymap, int m, std::vector<double>& out) {
      std::lock_guard<std::mutex> lock(xt_mutex);
      out.assign(static_cast<size_t>(m) * p, 0.0);
      double* ymap_dev = arena.alloc<double>(ymap.size(), false);
      double* out_dev = arena.alloc<double>(out.size(), false);
      CK(cudaMemcpyAsync(ymap_dev, ymap.data(), sizeof(double) * ymap.size(), cudaMemcpyHostToDevice, stream));
      if (raw.sparse)
        CK(sparse_xt_times(raw.csc_i, raw.csc_p, raw.csc_x, p, ddc.local, n_view, std_ ? ddc.x_scale : nullptr, ymap_dev, m, out_dev, stream));
      else
        CK(dense_xt_times(ddc.xd, n_view, p, ld, ymap_dev, m, out_dev, stream));
      CK(cudaMemcpyAsync(out.data(), out_dev, sizeof(double) * out.size(), cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
    };
    designs[key] = {hd, dd};
    return {hd, dd};
  }

  // ---------------------------------------------------------------------------------- one fit
  // Step 1 (serial): the design. Step 2 (may run on a helper thread, touches only the job): the plan. Step 3 (serial):
  // device state.
  std::string add_fit_design(const int32_t* rows, int64_t n_rows, const sgdnet_control& ctl, sgdnet_rng* rng,
                             const int32_t* test_rows, int64_t n_test, int32_t measure) {
    if (ctl.n_lambda <= 0) return "n_lambda must be positive";
    if (!rng) return "rng is null";
    if (rows)
      for (int64_t i = 0; i < n_rows; ++i)
        if (rows[i] < 0 || rows[i] >= raw.n) return "train row id out of range";
    FitJob job;
    auto dsn = get_design(rows, n_rows, ctl.standardize != 0);
    job.design = dsn.first;
    job.ddev = dsn.second;
    job.rng = rng;
    job.test_rows = test_rows;
    job.n_test = n_test;
    job.measure = measure;
    jobs.push_back(std::move(job));
    return "";
  }

  std::string build_plan(FitJob& job, const int32_t* rows, const sgdnet_control& ctl) {
    const HostDesign& d = *job.design;
    std::vector<double> ysub(static_cast<size_t>(d.n) * Ky);     // response restricted to the fit's rows
    for (int k = 0; k < Ky; ++k)
      for (int64_t i = 0; i < d.n; ++i)
        ysub[static_cast<size_t>(k) * d.n + i] = y_cm[static_cast<size_t>(k) * raw.n + (rows ? rows[i] : i)];
    return job.plan.build(d, std::move(ysub), Ky, ctl);
  }

  // device bytes alloc_fit is about to carve for this job (upper estimate), so that a batch reserves them in one go
  size_t fit_device_bytes(const FitJob& job) const {
    const HostDesign& d = *job.design;
    const FitPlan& pl = job.plan;
    const size_t K = pl.K, p = d.p, L = pl.n_lambda, n = d.n;
    const bool wave = d.sparse && K == 1 && !pl.standardize;
    size_t epl = std::max<int64_t>(1, std::min<int64_t>(16, 200000 / std::max<int64_t>(1, d.n)));
    if (const char* env = std::getenv("SGDNET_EPOCHS_PER_LAUNCH")) epl = std::max(1, std::atoi(env));
    size_t b = 3 * K * p * 8 + n * K * 8 + n * size_t(Ky) * 8 + p * 4 + (wave ? p * 32 : 0) + (d.sparse ? (n + 1) * 8 : 0);
    b += L * p * K * 8 + L * K * 8 + L * 24 + (pl.debug ? L * size_t(pl.max_iter) * 8 : 0);
    b += 2 * (epl * n * 4 + (wave ? epl * n * 257 : 0) + (epl + 1) * sizeof(MtState));
    b += size_t(sms) * 4 * 8 + sizeof(FitDev) + sizeof(Progress) + sizeof(MtState);
    return b + 64 * Arena::kAlign;
  }

  void alloc_fit(FitJob& job) {
    const HostDesign& d = *job.design;
    const FitPlan& pl = job.plan;
    const int K = pl.K, p = d.p, L = pl.n_lambda;
    FitDev& f = job.dev;
    f.sparse = d.sparse;
    f.family = pl.family;
    f.penalty = pl.penalty;
    f.fit_intercept = pl.fit_intercept;
    f.standardize = (d.sparse && pl.standardize) ? 1 : 0;
    f.K = K;
    f.Ky = Ky;
    f.p = p;
    f.ld = d.ld;
    f.n = d.n;
    f.xd = job.ddev.xd;
    f.rows = job.ddev.rows;
    f.ci = job.ddev.ci;
    f.cv = job.ddev.cv;
    f.c = job.ddev.c;
    f.yt = arena.upload(pl.yt);
    f.W = arena.alloc<double>(size_t(K) * p);
    f.gsum = arena.alloc<double>(size_t(K) * p);
    f.Wprev = arena.alloc<double>(size_t(K) * p);
    f.b = arena.upload(pl.intercept0);
    f.gsi = arena.alloc<double>(K);
    f.gmem = arena.alloc<double>(size_t(d.n) * K);
    f.lag = arena.alloc<uint32_t>(p);
    job.variant = !d.sparse ? Variant::Dense
                            : ((K == 1 && !f.standardize) ? Variant::SparseK1
                                                          : ((K == 1 && raw.max_row_nnz <= kCentCap && !no_centred) ? Variant::SparseCentred
                                                                                                                    : Variant::SparseGeneric));
    if (job.variant == Variant::SparseCentred) {
      bool in_smem = false;
      centred_smem_bytes(p, &in_smem);
      job.pos_scratch = in_smem ? nullptr : arena.alloc<uint16_t>(size_t(2) * p);
    }
    f.st = (job.variant == Variant::SparseK1) ? arena.alloc<FeatState>(p) : nullptr;
    f.lag_scaling = d.sparse ? arena.alloc<double>(size_t(d.n) + 1, false) : nullptr;
    f.gamma = arena.upload(pl.gamma);
    f.alpha = arena.upload(pl.alpha);
    f.beta = arena.upload(pl.beta);
    f.n_lambda = L;
    f.max_iter = pl.max_iter;
    f.tol = pl.tol;
    f.null_deviance_scaled = pl.nulldev_scaled;
    f.x_center = job.ddev.x_center;
    f.x_scale = job.ddev.x_scale;
    f.y_center = arena.upload(pl.y_center);
    f.y_scale = arena.upload(pl.y_scale);
    f.beta_arch = arena.alloc<double>(size_t(L) * p * K, false);
    f.a0_arch = arena.alloc<double>(size_t(L) * K, false);
    f.dev_ratio = arena.alloc<double>(L, false);
    f.epochs = arena.alloc<uint32_t>(L);
    f.codes = arena.alloc<uint32_t>(L);
    f.debug = pl.debug ? 1 : 0;
    f.losses = pl.debug ? arena.alloc<double>(size_t(L) * pl.max_iter) : nullptr;

    // epochs per launch: amortise the host visit for small problems; the callback generator and the debug loss need
    // one after every epoch
    int epl = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(16, 200000 / std::max<int64_t>(1, d.n))));
    if (const char* env = std::getenv("SGDNET_EPOCHS_PER_LAUNCH")) epl = std::max(1, std::atoi(env));   // tuning knob
    if (pl.debug || job.rng->kind == SGDNET_RNG_CALLBACK) epl = 1;
    epl = static_cast<int>(std::min<uint32_t>(static_cast<uint32_t>(epl), std::max<uint32_t>(1u, pl.max_iter)));
    job.epochs_per_launch = epl;
    job.device_rng = job.rng->kind == SGDNET_RNG_MT && !host_rng;
    const int nbuf = job.device_rng ? 2 : 1;
    for (int b = 0; b < nbuf; ++b) {
      job.seq_dev[b] = arena.alloc<uint32_t>(size_t(epl) * d.n, false);
      if (job.variant == Variant::SparseK1) {
        job.dep_dev[b] = arena.alloc<uint64_t>(size_t(epl) * d.n * 32, false);
        job.dup_dev[b] = arena.alloc<uint8_t>(size_t(epl) * d.n, false);
      }
      if (job.device_rng) job.snap_dev[b] = arena.alloc<MtState>(size_t(epl) + 1, false);
    }
    if (job.device_rng) {
      job.rng_dev = arena.alloc<MtState>(1, false);
      job.rng_pin = arena.host<MtState>(1);
    } else {
      job.seq_pin = arena.host<uint32_t>(size_t(epl) * d.n);
    }
    // streaming passes: enough CTAs to fill the GPU when the fit's pass runs alone
    job.loss_blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(int64_t(sms) * 4, (d.n + 7) / 8)));
    f.partials = arena.alloc<double>(job.loss_blocks);
    f.xb_partials = arena.alloc<double>(size_t(kRescaleBlocks) * K);
    const int mask_words = (p + 31) / 32;
    // sparse, K == 1, no virtual centring: the bulk-copy tile form of the loss pass, one CTA per SM
    job.loss_tiles = job.variant == Variant::SparseK1
                         ? static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(sms, job.loss_blocks), (d.n + 31) / 32)))
                         : 0;
    if (no_loss_tiles) job.loss_tiles = 0;
    job.mask_words = (job.variant == Variant::SparseK1 && mask_words <= loss_mask_words_max(job.loss_tiles > 0)) ? mask_words : 0;
    f.nz_mask = job.mask_words ? arena.alloc<uint32_t>(mask_words) : nullptr;
    job.dense_cluster = job.variant == Variant::Dense && p >= SGD_WIDE_P;
    if (job.dense_cluster) {
      job.dense_smem = dense_cluster_smem_bytes(K, p, f.penalty);
      if (job.dense_smem > dense_smem_budget())
        throw std::invalid_argument("dense x with p = " + std::to_string(p) + " columns: the row ring of one cluster CTA exceeds its shared memory");
    } else if (job.variant == Variant::Dense) {
      int in_smem = 0;
      job.dense_smem = dense_smem_bytes(K, p, d.ld, &in_smem);
      if (job.dense_smem > dense_smem_budget())
        throw std::invalid_argument("dense x with p = " + std::to_string(p) + " columns: a row ring of " +
                                    std::to_string(job.dense_smem) + " bytes exceeds one SM's shared memory (" +
                                    std::to_string(dense_smem_budget()) + "); this build handles dense p up to about 7000");
    }
    job.mirror = arena.host<Progress>(1);
    std::memset(job.mirror, 0, sizeof(Progress));
    job.mirror->wscale = 1.0;
    f.mirror = job.mirror;
    job.dev_ptr = arena.upload_from(&job.dev, 1);
    job.prog_ptr = arena.upload_from(job.mirror, 1);
    CK(cudaStreamCreateWithFlags(&job.st, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&job.st_prep, cudaStreamNonBlocking));
    if (f.st != nullptr && l2_persist_bytes > 0) {
      // the packed coefficient records are gathered and scattered once per nonzero per update while 1.2 KB of row
      // data per update streams through L2: ask L2 to keep the records (ncu: a quarter of their sectors came from HBM)
      cudaStreamAttrValue av{};
      av.accessPolicyWindow.base_ptr = f.st;
      av.accessPolicyWindow.num_bytes = std::min<size_t>(size_t(p) * sizeof(FeatState), l2_window_max);
      av.accessPolicyWindow.hitRatio = 1.0f;
      av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      (void)cudaStreamSetAttribute(job.st, cudaStreamAttributeAccessPolicyWindow, &av);   // a hint: failure is not an error
      (void)cudaGetLastError();
    }
    for (cudaEvent_t* ev : {&job.ev0, &job.ev1, &job.ev_f0, &job.ev_f1}) CK(cudaEventCreate(ev));
    for (cudaEvent_t* ev : {&job.ev_idx, &job.ev_prep}) CK(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
  }

  // add_fit_design + build_plan + alloc_fit for one fit (single-fit entry points)
  std::string add_fit(const int32_t* rows, int64_t n_rows, const sgdnet_control& ctl, sgdnet_rng* rng,
                      const int32_t* test_rows, int64_t n_test) {
    std::string err = add_fit_design(rows, n_rows, ctl, rng, test_rows, n_test, 0);
    if (!err.empty()) return err;
    PhaseTimer pt;
    err = build_plan(jobs.back(), rows, ctl);
    if (!err.empty()) {
      jobs.pop_back();
      return err;
    }
    pt.lap("plan (lambda path, steps)");
    alloc_fit(jobs.back());
    pt.lap("state alloc + upload");
    return "";
  }

  void finalize_batch() {
    // Everything uploaded so far went through cudaMemcpy from pageable memory, which may return while the DMA into
    // device memory is still in flight, on the legacy default stream - and the fits' streams are non-blocking, so
    // nothing orders their kernels behind it. Zero fills are ordered on `stream`. One device-wide wait covers both.
    CK(cudaDeviceSynchronize());
    seconds_setup = now_s() - t_begin;
  }

  // ---------------------------------------------------------------------------------- index stream (host generators)
  // Makes sure `need` undrawn-by-the-device indices are pending (host memory).
  bool ensure_pending(FitJob& j, size_t need) {
    size_t have = j.pending.size() - j.pending_head;
    if (have >= need) return true;
    if (j.pending_head > 0) {
      j.pending.erase(j.pending.begin(), j.pending.begin() + j.pending_head);
      j.pending_head = 0;
    }
    const size_t add = need - have;
    // keep every mark from the newest one that is not ahead of the consumed position: settle_rng restarts from it
    size_t keep_from = 0;
    for (size_t i = 0; i < j.marks.size(); ++i)
      if (j.marks[i].first <= j.consumed) keep_from = i;
    if (keep_from > 0) j.marks.erase(j.marks.begin(), j.marks.begin() + keep_from);
    j.marks.emplace_back(j.generated, *j.rng);
    const size_t old = j.pending.size();
    j.pending.resize(old + add);
    if (!draw_indices(j.rng, static_cast<uint32_t>(j.dev.n), static_cast<int64_t>(add), j.pending.data() + old)) return false;
    j.generated += add;
    return true;
  }

  bool stage_indices(FitJob& j, int n_epochs) {
    const size_t need = size_t(n_epochs) * j.dev.n;
    if (!ensure_pending(j, need)) return false;
    std::memcpy(j.seq_pin, j.pending.data() + j.pending_head, need * sizeof(uint32_t));
    return true;
  }

  // Give the caller's generator back advanced by exactly the draws the fit consumed.
  void settle_rng(FitJob& j) {
    if (j.device_rng) {
      // the state the next launch would have started from: snapshot `epochs used` of the last launch
      if (j.cur_state == nullptr) return;
      CK(cudaMemcpyAsync(j.rng_pin, j.cur_state, sizeof(MtState), cudaMemcpyDeviceToHost, j.st));
      CK(cudaStreamSynchronize(j.st));
      std::memcpy(j.rng->mt, j.rng_pin->mt, sizeof(j.rng->mt));
      j.rng->mti = j.rng_pin->mti;
      j.handed_back = *j.rng;
      j.have_handed_back = true;
      return;
    }
    if (j.generated == j.consumed) return;
    if (j.rng->kind == SGDNET_RNG_SEQUENCE) {
      j.rng->seq_pos -= static_cast<int64_t>(j.generated - j.consumed);
      j.generated = j.consumed;
      j.pending.clear();
      j.pending_head = 0;
      return;
    }
    if (j.rng->kind == SGDNET_RNG_CALLBACK) return;   // epochs_per_launch == 1: nothing was drawn ahead
    for (int i = static_cast<int>(j.marks.size()) - 1; i >= 0; --i) {
      if (j.marks[i].first <= j.consumed) {
        const sgdnet_rng keep = *j.rng;
        *j.rng = j.marks[i].second;
        j.rng->unif_rand = keep.unif_rand;
        j.rng->ctx = keep.ctx;
        for (uint64_t q = j.marks[i].first; q < j.consumed; ++q) (void)mt_unif(j.rng);
        j.generated = j.consumed;
        j.pending.clear();
        j.pending_head = 0;
        j.marks.clear();
        return;
      }
    }
    throw std::runtime_error("internal: no generator mark at or before the consumed position");
  }

  // ---------------------------------------------------------------------------------- launches of one fit
  // CTAs one fit's conflict-code pass may use: the GPU shared among the fits in flight
  int deps_ctas() const {
    int active = 0;
    for (const FitJob& j : jobs) active += (j.phase != Phase::Done && j.phase != Phase::Parked) ? 1 : 0;
    if (deps_ctas_env > 0) return deps_ctas_env;
    return std::max(8, sms * 6 / std::max(1, active));
  }
  int deps_ctas_env = std::getenv("SGDNET_DEPS_CTAS") ? std::atoi(std::getenv("SGDNET_DEPS_CTAS")) : 0;   // measurement aid

  // the caller's generator -> device, at the start of run(); a launch prepared ahead stays valid when the caller hands
  // back the generator exactly as it received it
  void upload_rng(FitJob& j) {
    if (!j.device_rng) return;
    const bool same = j.have_handed_back && j.rng->mti == j.handed_back.mti &&
                      std::memcmp(j.rng->mt, j.handed_back.mt, sizeof(j.rng->mt)) == 0;
    if (same && j.cur_state != nullptr) return;
    CK(cudaStreamSynchronize(j.st_prep));
    std::memcpy(j.rng_pin->mt, j.rng->mt, sizeof(j.rng->mt));
    j.rng_pin->mti = j.rng->mti;
    CK(cudaMemcpyAsync(j.rng_dev, j.rng_pin, sizeof(MtState), cudaMemcpyHostToDevice, j.st));
    j.cur_state = j.rng_dev;
    j.prepped = false;
    j.prepped_next = false;
  }

  // `ne_fixed` > 0: measurement mode, run exactly that many epochs (<= epochs_per_launch)
  void submit_solver(FitJob& j, int flags, int ne_fixed = 0) {
    const Progress pg = *j.mirror;
    const uint32_t left = j.plan.max_iter - pg.it_outer;
    const int epl = j.epochs_per_launch;
    const int ne = ne_fixed > 0 ? ne_fixed
                                : static_cast<int>(std::min<uint32_t>(static_cast<uint32_t>(epl), std::max<uint32_t>(left, 1u)));
    const int b = j.buf;
    {      // NVTX: one range per lambda and one per solver launch of every fit (they overlap across fits: start/end ranges)
      char name[96];
      const int fit_no = static_cast<int>(&j - jobs.data());
      if (!j.nvtx_lambda_open) {
        std::snprintf(name, sizeof(name), "sgdnet fit %d lambda %d", fit_no, pg.lambda_ind);
        j.nvtx_lambda = nvtxRangeStartA(name);
        j.nvtx_lambda_open = true;
      }
      std::snprintf(name, sizeof(name), "sgdnet fit %d lambda %d launch %u (%d epochs)", fit_no, pg.lambda_ind, j.round_id + 1, ne);
      j.nvtx_launch = nvtxRangeStartA(name);
    }
    RoundArgs ra{j.seq_dev[b], j.dep_dev[b], j.dup_dev[b], ne, flags, ++j.round_id, 0u};
    const int64_t n = j.dev.n;
    if (j.prepped) {
      CK(cudaStreamWaitEvent(j.st, j.ev_prep, 0));
      j.idx_on_prep = true;
    } else {
      if (j.stale_prep) {
        CK(cudaStreamWaitEvent(j.st, j.ev_prep, 0));   // the discarded preparation reads what is rewritten below
        j.stale_prep = false;
      }
      if (j.device_rng) {
        CK(launch_mt_indices(j.cur_state, static_cast<uint32_t>(n), epl, j.seq_dev[b], j.snap_dev[b], j.st));
        ++j.launches;
      } else {
        if (!stage_indices(j, ne)) {
          g_error = "sampling-index source exhausted";
          throw CudaFail{cudaSuccess, "rng"};
        }
        CK(cudaMemcpyAsync(j.seq_dev[b], j.seq_pin, size_t(ne) * n * sizeof(uint32_t), cudaMemcpyHostToDevice, j.st));
      }
      CK(cudaEventRecord(j.ev_idx, j.st));
      j.idx_on_prep = false;
      if (j.variant == Variant::SparseK1) {
        RoundArgs rd = ra;
        rd.n_epochs = j.device_rng ? epl : ne;
        CK(launch_wave_deps(j.dev_ptr, rd, n * rd.n_epochs, deps_ctas(), j.st));
        ++j.launches;
      }
    }
    if (j.variant != Variant::Dense) {
      CK(launch_lag_scaling(j.dev_ptr, j.prog_ptr, j.st));
      ++j.launches;
    }
    // CUDA events bracket the solver only when the fit has the GPU to itself (single fits, the stepping interface): an
    // event recorded behind a long kernel holds up the hardware queue its stream shares with other fits' streams (at
    // most 32 queues), which caps how many fits of a batch run at once. Batches use the kernels' own %globaltimer
    // brackets (Progress::solver_ns).
    if (use_events()) CK(cudaEventRecord(j.ev0, j.st));
    if (j.dense_cluster)
      CK(launch_saga_dense_cluster(j.dev.K, j.dev.p, j.dev.penalty, j.dense_smem, j.dev_ptr, j.prog_ptr, ra, j.st));
    else if (j.variant == Variant::Dense)
      CK(launch_saga_dense(j.dev.K, j.dev.penalty, j.dense_smem, j.dev_ptr, j.prog_ptr, ra, j.st));
    else if (j.variant == Variant::SparseCentred) {
      // the instantiation for this lambda's penalty: r = 1 - alpha gamma == 1 (the lasso) keeps wscale at exactly 1
      const int li = std::min<int>(pg.lambda_ind, static_cast<int>(j.plan.alpha.size()) - 1);
      const bool ident = (1.0 - j.plan.alpha[li] * j.plan.gamma[li]) == 1.0;
      const int mode = j.dev.penalty == kElasticNet ? (ident ? 0 : 1) : ((j.dev.penalty == kRidge && !ident) ? 2 : 3);
      CK(launch_saga_sparse_centred(j.dev.p, mode, j.dev_ptr, j.prog_ptr, ra, j.pos_scratch, j.st));
    }
    else
      CK(launch_saga_sparse(j.variant == Variant::SparseK1, j.dev_ptr, j.prog_ptr, ra, j.st));
    ++j.launches;
    if (use_events()) CK(cudaEventRecord(j.ev1, j.st));
    j.ne_submitted = ne;
    j.t_submit = now_s();
    j.prepped = false;
    j.prepped_next = false;
    // ---- the next launch's indices and conflict codes, prepared on the second stream while this one runs, on the
    // assumption that this launch consumes all `ne` epochs (it does unless the lambda converges inside it)
    if (j.device_rng && !no_overlap) {
      if (!j.idx_on_prep) CK(cudaStreamWaitEvent(j.st_prep, j.ev_idx, 0));
      const int nb = b ^ 1;
      CK(launch_mt_indices(j.snap_dev[b] + ne, static_cast<uint32_t>(n), epl, j.seq_dev[nb], j.snap_dev[nb], j.st_prep));
      ++j.launches;
      if (j.variant == Variant::SparseK1) {
        RoundArgs rd{j.seq_dev[nb], j.dep_dev[nb], j.dup_dev[nb], epl, 0, 0u, 0u};
        CK(launch_wave_deps(j.dev_ptr, rd, n * epl, deps_ctas(), j.st_prep));
        ++j.launches;
      }
      CK(cudaEventRecord(j.ev_prep, j.st_prep));
      j.prepped_next = true;
    }
    j.phase = Phase::Solver;
  }

  void solver_done(FitJob& j) {
    nvtxRangeEnd(j.nvtx_launch);
    const Progress pg = *j.mirror;
    const uint64_t used_epochs = pg.epochs_last_launch;
    const uint64_t used = used_epochs * uint64_t(j.dev.n);
    float ms = static_cast<float>((pg.solver_ns - j.solver_ns_seen) * 1e-6);
    j.solver_ns_seen = pg.solver_ns;
    if (use_events()) {
      CK(cudaEventSynchronize(j.ev1));
      CK(cudaEventElapsedTime(&ms, j.ev0, j.ev1));
    }
    j.seconds_solver += ms * 1e-3;
    if (trace_rounds)
      std::fprintf(stderr, "[sgdnet_b200] t=%9.3f ms fit %d: launch %u done (submitted t=%9.3f), %d epochs at lambda %d, solver %.3f ms%s\n",
                   (now_s() - t_run) * 1e3, static_cast<int>(&j - jobs.data()), j.round_id, (j.t_submit - t_run) * 1e3,
                   static_cast<int>(used_epochs), pg.lambda_ind, ms, j.idx_on_prep ? " (prepared ahead)" : "");
    if (j.device_rng) {
      j.cur_state = j.snap_dev[j.buf] + used_epochs;
      if (j.prepped_next && used_epochs == static_cast<uint64_t>(j.ne_submitted)) {
        j.buf ^= 1;
        j.prepped = true;
      } else if (j.prepped_next) {
        j.stale_prep = true;
      }
      j.prepped_next = false;
    } else {
      j.pending_head += used;
      j.consumed += used;
    }
    j.needs_finish = pg.status == kLambdaDone || (j.dev.debug && used_epochs > 0);
  }

  void submit_finish(FitJob& j) {
    ++j.round_id;
    j.t_finish = now_s();
    if (use_events()) CK(cudaEventRecord(j.ev_f0, j.st));
    if (j.dev.debug) {
      CK(launch_epoch_loss(j.dev_ptr, j.prog_ptr, j.loss_blocks, j.loss_tiles, j.st));
      j.launches += 2;
    }
    CK(launch_finish_lambda(j.dev_ptr, j.prog_ptr, j.loss_blocks, j.loss_tiles, j.mask_words, j.round_id, j.st));
    j.launches += 3;
    if (use_events()) CK(cudaEventRecord(j.ev_f1, j.st));
    j.phase = Phase::Finish;
  }

  void finish_done(FitJob& j) {
    float ms = static_cast<float>((now_s() - j.t_finish) * 1e3);      // batches: host clock from submission to publication
    if (use_events()) {
      CK(cudaEventSynchronize(j.ev_f1));
      CK(cudaEventElapsedTime(&ms, j.ev_f0, j.ev_f1));
    }
    j.seconds_dev += ms * 1e-3;
    j.needs_finish = false;
    if (j.nvtx_lambda_open && j.mirror->status != kLambdaDone) {      // the lambda's deviance / archive pass is through
      nvtxRangeEnd(j.nvtx_lambda);
      j.nvtx_lambda_open = false;
    }
  }

  bool use_events() const { return jobs.size() == 1; }
  uint64_t idle_sweeps = 0;
  void wait_round(FitJob& j) {
    uint64_t spins = 0;
    while (*reinterpret_cast<volatile uint32_t*>(&j.mirror->round_seen) != j.round_id) {
      if (++spins % 4096 == 0) check_in_flight();
      std::this_thread::yield();
    }
    std::atomic_thread_fence(std::memory_order_acquire);
  }
  void check_in_flight() {
    for (FitJob& j : jobs) {
      if (j.phase != Phase::Solver && j.phase != Phase::Finish) continue;
      const cudaError_t e = cudaStreamQuery(j.st);
      if (e == cudaErrorNotReady) continue;
      if (e != cudaSuccess) throw CudaFail{e, "a launch of the batch failed"};
      // the stream has drained: the round must have been published
      std::atomic_thread_fence(std::memory_order_acquire);
      if (*reinterpret_cast<volatile uint32_t*>(&j.mirror->round_seen) != j.round_id)
        throw std::runtime_error("internal: a launch finished without publishing its progress");
    }
  }

  // ---------------------------------------------------------------------------------- the event loop
  // Runs every job to the end of its path (or, with `only_lambda` >= 0, until that lambda is finished). `score`: fits
  // with held-out rows are scored on their own stream as soon as they are done.
  void run(int only_lambda = -1, bool score = false) {
    t_run = now_s();
    for (FitJob& j : jobs) {
      if (j.phase == Phase::Parked) j.phase = Phase::Idle;
      if (j.phase != Phase::Done) upload_rng(j);
    }
    for (;;) {
      bool progressed = false, any_live = false;
      for (FitJob& j : jobs) {
        if (j.phase == Phase::Done || j.phase == Phase::Parked) continue;
        any_live = true;
        if (j.phase == Phase::Solver || j.phase == Phase::Finish) {
          if (*reinterpret_cast<volatile uint32_t*>(&j.mirror->round_seen) != j.round_id) continue;
          std::atomic_thread_fence(std::memory_order_acquire);
          if (j.phase == Phase::Solver) solver_done(j);
          else finish_done(j);
          j.phase = Phase::Idle;
        }
        // Idle: decide the fit's next launch from its published progress
        const Progress pg = *j.mirror;
        // A launch whose indices and conflict codes were prepared ahead is submitted only once that preparation has
        // finished (host-side query). Waiting for it inside the stream instead (cudaStreamWaitEvent) would park a blocked
        // wait at the head of one of the 32 hardware queues the batch's streams share, and every other fit's kernels
        // queued behind it would stall with it.
        if (jobs.size() > 1 && j.prepped && pg.status == kRunning && !(j.needs_finish || pg.status == kLambdaDone) &&
            cudaEventQuery(j.ev_prep) == cudaErrorNotReady)
          continue;
        progressed = true;
        if (pg.status == kFitDone) {
          if (score && j.test_rows && j.n_test > 0 && !j.scored) submit_score(j);
          j.phase = Phase::Done;
        } else if (j.needs_finish || pg.status == kLambdaDone) {
          submit_finish(j);
        } else if (only_lambda >= 0 && pg.lambda_ind > only_lambda) {
          j.phase = Phase::Parked;
        } else {
          submit_solver(j, 0);
        }
      }
      if (!any_live) break;
      if (!progressed) {
        // nothing finished in this sweep: let the core breathe; now and then make sure the launches in flight are
        // still alive (a failed launch never publishes its round)
        if (++idle_sweeps % 4096 == 0) check_in_flight();
        std::this_thread::yield();
      } else {
        idle_sweeps = 0;
      }
    }
    for (FitJob& j : jobs) {
      if (j.path_only) continue;
      CK(cudaStreamSynchronize(j.st));
      CK(cudaStreamSynchronize(j.st_prep));
    }
    if (trace_rounds) std::fprintf(stderr, "[sgdnet_b200] run: %.3f ms\n", (now_s() - t_run) * 1e3);
  }

  // ---------------------------------------------------------------------------------- results
  void fill_result(int i, sgdnet_result* out) {
    FitJob& j = jobs[i];
    const FitPlan& pl = j.plan;
    const int L = pl.n_lambda, K = pl.K, p = j.design->p;
    std::memset(out, 0, sizeof(*out));
    out->n_lambda = L;
    out->n_classes = K;
    out->n_features = p;
    auto mal = [](size_t bytes) { return std::malloc(std::max<size_t>(bytes, 8)); };
    out->a0 = static_cast<double*>(mal(sizeof(double) * L * K));
    out->beta = static_cast<double*>(mal(sizeof(double) * size_t(L) * p * K));
    out->lambda = static_cast<double*>(mal(sizeof(double) * L));
    out->dev_ratio = static_cast<double*>(mal(sizeof(double) * L));
    out->return_codes = static_cast<uint32_t*>(mal(sizeof(uint32_t) * L));
    out->epochs = static_cast<uint32_t*>(mal(sizeof(uint32_t) * L));
    out->losses_ptr = static_cast<int64_t*>(mal(sizeof(int64_t) * (L + 1)));
    if (!out->a0 || !out->beta || !out->lambda || !out->dev_ratio || !out->return_codes || !out->epochs || !out->losses_ptr)
      throw CudaFail{cudaErrorMemoryAllocation, "result buffers"};
    std::memcpy(out->lambda, pl.lambda.data(), sizeof(double) * L);
    out->nulldev = pl.nulldev;
    if (j.path_only) {               // the lambda path and the null deviance are all there is
      std::memset(out->a0, 0, sizeof(double) * L * K);
      std::memset(out->beta, 0, sizeof(double) * size_t(L) * p * K);
      std::memset(out->dev_ratio, 0, sizeof(double) * L);
      std::memset(out->return_codes, 0, sizeof(uint32_t) * L);
      std::memset(out->epochs, 0, sizeof(uint32_t) * L);
      out->losses = static_cast<double*>(mal(8));
      for (int l = 0; l <= L; ++l) out->losses_ptr[l] = 0;
      out->seconds_setup = seconds_setup;
      out->seconds_total = now_s() - t_begin;
      return;
    }
    CK(cudaMemcpy(out->a0, j.dev.a0_arch, sizeof(double) * L * K, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out->beta, j.dev.beta_arch, sizeof(double) * size_t(L) * p * K, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out->dev_ratio, j.dev.dev_ratio, sizeof(double) * L, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out->return_codes, j.dev.codes, sizeof(uint32_t) * L, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out->epochs, j.dev.epochs, sizeof(uint32_t) * L, cudaMemcpyDeviceToHost));
    uint32_t np = 0;
    for (int l = 0; l < L; ++l) np += out->epochs[l];
    out->npasses = np;
    out->losses_ptr[0] = 0;
    if (pl.debug) {
      std::vector<double> all(size_t(L) * pl.max_iter);
      CK(cudaMemcpy(all.data(), j.dev.losses, sizeof(double) * all.size(), cudaMemcpyDeviceToHost));
      out->losses = static_cast<double*>(mal(sizeof(double) * np));
      size_t w = 0;
      for (int l = 0; l < L; ++l) {
        for (uint32_t e = 0; e < out->epochs[l]; ++e) out->losses[w++] = all[size_t(l) * pl.max_iter + e];
        out->losses_ptr[l + 1] = static_cast<int64_t>(w);
      }
    } else {
      out->losses = static_cast<double*>(mal(8));
      for (int l = 0; l < L; ++l) out->losses_ptr[l + 1] = 0;
    }
    out->seconds_solver = j.seconds_solver;
    out->seconds_deviance = j.seconds_dev;
    out->seconds_setup = seconds_setup;
    out->seconds_total = now_s() - t_begin;
    out->kernel_launches = j.launches;
  }

  // ---------------------------------------------------------------------------------- raw design + scoring
  void upload_raw() {
    if (raw_uploaded) return;
    auto dsn = get_design(nullptr, raw.n, false);
    raw_dev = dsn.second;
    std::vector<double> yt(static_cast<size_t>(raw.n) * Ky);
    for (int k = 0; k < Ky; ++k)
      for (int64_t i = 0; i < raw.n; ++i) yt[static_cast<size_t>(i) * Ky + k] = y_cm.empty() ? 0.0 : y_cm[static_cast<size_t>(k) * raw.n + i];
    yraw_dev = arena.upload(yt);
    raw_uploaded = true;
  }

  // score/link of rows `row_ids` (device pointer or null) under coefficients (a0_dev, beta_dev), on stream `st`.
  // `measure`: SGDNET_MEASURE_* (R/score.R:55-178). Multinomial "class" may need a second pass (see the kernel): the
  // call then waits for the first one on `st`.
  uint64_t predict_score(int family, int K, int L, const int32_t* row_ids_dev, int64_t n_rows, const double* a0_dev,
                         const double* beta_dev, bool with_y, double* link_dev, double* score_dev, cudaStream_t st,
                         int measure = SGDNET_MEASURE_DEVIANCE) {
    upload_raw();
    const HostDesign& hd = *designs[std::make_pair((const int32_t*)nullptr, 0)].first;
    PredictArgs a{};
    a.sparse = hd.sparse;
    a.family = family;
    a.K = K;
    a.Ky = Ky;
    a.p = hd.p;
    a.ld = hd.ld;
    a.n_lambda = L;
    a.measure = measure;
    a.n = n_rows;
    a.row_ids = row_ids_dev;
    a.xd = raw_dev.xd;
    a.rows = raw_dev.rows;
    a.ci = raw_dev.ci;
    a.cv = raw_dev.cv;
    a.y = with_y ? yraw_dev : nullptr;
    a.a0 = a0_dev;
    a.beta = beta_dev;
    a.link = link_dev;
    const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(int64_t(sms) * 4, (n_rows + 7) / 8)));
    a.partials = arena.alloc<double>(size_t(blocks) * L, false);     // every entry is written by its block
    a.score = score_dev;
    double* bt = arena.alloc<double>(size_t(hd.p) * L * K, false);
    uint64_t launches = score_dev ? 3 : 2;
    const bool two_pass = with_y && family == kMultinomial && measure == SGDNET_MEASURE_CLASS;
    uint32_t* present_dev = nullptr;
    if (two_pass) {
      present_dev = arena.alloc<uint32_t>(1, false);
      CK(cudaMemsetAsync(present_dev, 0, sizeof(uint32_t), st));
      a.present = present_dev;
    }
    CK(launch_predict_score(a, bt, blocks, st, true));
    if (two_pass) {
      uint32_t present = 0;
      CK(cudaMemcpyAsync(&present, present_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      const uint32_t all = (K >= 32) ? 0xffffffffu : ((1u << K) - 1u);
      if ((present & all) != all) {
        // some class is never predicted: as.numeric(as.factor(.)) numbers the predicted classes 1, 2, ... in order
        std::vector<int32_t> remap(K, 0);
        int32_t rank = 0;
        for (int k = 0; k < K; ++k) {
          remap[k] = rank;
          if (present & (1u << k)) ++rank;
        }
        int32_t* remap_dev = arena.alloc<int32_t>(K, false);
        CK(cudaMemcpyAsync(remap_dev, remap.data(), sizeof(int32_t) * K, cudaMemcpyHostToDevice, st));
        a.remap = remap_dev;
        a.present = nullptr;
        CK(launch_predict_score(a, bt, blocks, st, false));
        CK(cudaStreamSynchronize(st));       // `remap` lives on this stack frame
        launches += 2;
      }
    }
    return launches;
  }

  // held-out rows and the raw design / response on the device, before the batch starts (see finalize_batch)
  void upload_for_scoring() {
    bool any = false;
    for (FitJob& j : jobs) {
      if (j.path_only || !j.test_rows || j.n_test <= 0) continue;
      any = true;
      int32_t*& td = test_dev[j.test_rows];
      if (!td) {
        td = arena.alloc<int32_t>(j.n_test, false);
        CK(cudaMemcpy(td, j.test_rows, sizeof(int32_t) * j.n_test, cudaMemcpyHostToDevice));
      }
    }
    if (any) upload_raw();
  }

  void submit_score(FitJob& j) {
    int32_t* td = test_dev[j.test_rows];
    const int L = j.plan.n_lambda;
    j.score_dev = arena.alloc<double>(L, false);
    j.launches += predict_score(j.plan.family, j.plan.K, L, td, j.n_test, j.dev.a0_arch, j.dev.beta_arch, true, nullptr,
                                j.score_dev, j.st, j.measure);
    j.scored = true;
  }
};

// ============================================================================================ C ABI helpers
int fail(int code, const std::string& msg) {
  g_error = msg;
  return code;
}

template <typename F>
int guarded(F&& body) {
  try {
    return body();
  } catch (const CudaFail& f) {
    if (f.e != cudaSuccess) g_error = std::string(f.what) + ": " + cudaGetErrorString(f.e);
    else if (g_error.empty()) g_error = f.what;
    return (f.e == cudaSuccess) ? SGDNET_ERR_RNG : SGDNET_ERR_CUDA;
  } catch (const std::invalid_argument& e) {
    g_error = e.what();
    return SGDNET_ERR_ARG;
  } catch (const std::bad_alloc&) {
    g_error = "host allocation failed";
    return SGDNET_ERR_ALLOC;
  } catch (const std::exception& e) {
    g_error = e.what();
    return SGDNET_ERR_INTERNAL;
  }
}

bool basic_args(int64_t n, int64_t p, const double* y, int32_t y_cols, std::string& why) {
  if (n <= 0 || p <= 0) { why = "x must have at least one row and one column"; return false; }
  if (n > 0xffffffffLL || p > 0x7fffffffLL) { why = "n and p must fit 32-bit indices"; return false; }
  if (!y || y_cols <= 0) { why = "y is null or has no columns"; return false; }
  return true;
}

struct XArg {
  bool sparse;
  const double* x;
  const int32_t *ci, *cp;
  const double* cx;
  int64_t n, p;
};

void load_x(Engine& eng, const XArg& xa, const double* y, int32_t y_cols) {
  PhaseTimer pt;
  if (xa.sparse) eng.load_sparse(xa.ci, xa.cp, xa.cx, xa.n, xa.p);
  else eng.load_dense(xa.x, xa.n, xa.p);
  eng.Ky = y_cols;
  if (y) eng.y_cm.assign(y, y + static_cast<size_t>(xa.n) * y_cols);
}

int fit_single(const XArg& xa, const double* y, int32_t y_cols, const sgdnet_control* control, sgdnet_rng* rng,
               sgdnet_result* out) {
  std::string why;
  if (!control || !rng || !out) return fail(SGDNET_ERR_ARG, "null control, rng or result");
  if (!basic_args(xa.n, xa.p, y, y_cols, why)) return fail(SGDNET_ERR_ARG, why);
  return guarded([&]() -> int {
    Engine eng;
    load_x(eng, xa, y, y_cols);
    std::string err = eng.add_fit(nullptr, xa.n, *control, rng, nullptr, 0);
    if (!err.empty()) return fail(SGDNET_ERR_ARG, err);
    eng.finalize_batch();
    eng.run();
    eng.settle_rng(eng.jobs[0]);
    eng.fill_result(0, out);
    return SGDNET_OK;
  });
}

bool measure_ok(int family, int measure, std::string& why) {
  // the measures each family's score() accepts (R/score.R:58, 78-82, 124-127, 165)
  const bool binomial = family == SGDNET_BINOMIAL, multinomial = family == SGDNET_MULTINOMIAL;
  if (measure < SGDNET_MEASURE_DEVIANCE || measure > SGDNET_MEASURE_AUC) { why = "unknown type.measure"; return false; }
  if (measure == SGDNET_MEASURE_CLASS && !(binomial || multinomial)) { why = "type.measure 'class' needs a binomial or multinomial fit"; return false; }
  if (measure == SGDNET_MEASURE_AUC && !binomial) { why = "type.measure 'auc' needs a binomial fit"; return false; }
  return true;
}

// auc(y, prob) of R/score.R:203-232 for one lambda, from the probabilities the device computed. The reference doubles
// the data (every observation once as a 0 with weight y1, once as a 1 with weight y2), breaks ties in `prob` with
// stats::runif - one draw per doubled observation, 2n per lambda, taken here from the caller's generator on the calling
// thread - and sums, over the observations of the second class, the weight of the first-class observations ordered
// before them. Only the entries with weight 1 matter: first-class observation i carries draw r[i], second-class
// observation i draw r[n + i]; on a full tie the first-class entry comes first (it has the smaller index).
double auc_one_lambda(const double* eta, const double* y, int64_t n, sgdnet_rng* rng) {
  std::vector<double> r(static_cast<size_t>(2 * n));
  for (double& v : r) v = (rng->kind == SGDNET_RNG_CALLBACK) ? rng->unif_rand(rng->ctx) : mt_unif(rng);
  struct Item {
    double prob, r;
    int32_t pos;      // 0: first class (y == 0), 1: second class
  };
  std::vector<Item> items(static_cast<size_t>(n));
  for (int64_t i = 0; i < n; ++i) {
    const bool second = y[i] > 0.5;
    items[i] = Item{1.0 / (1.0 + std::exp(-eta[i])), second ? r[n + i] : r[i], second ? 1 : 0};
  }
  std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) {
    if (a.prob != b.prob) return a.prob < b.prob;
    if (a.r != b.r) return a.r < b.r;
    return a.pos < b.pos;
  });
  double first_seen = 0.0, u = 0.0, n1 = 0.0;
  for (const Item& it : items) {
    if (it.pos) {
      u += first_seen;
      n1 += 1.0;
    } else {
      first_seen += 1.0;
    }
  }
  const double n0 = static_cast<double>(n) - n1;
  return std::exp(std::log(u) - std::log(n1) - std::log(n0));
}

int fit_batch(const XArg& xa, const double* y, int32_t y_cols, sgdnet_fit_spec* specs, int32_t n_fits,
              sgdnet_result* results, double* scores) {
  std::string why;
  if (!specs || n_fits <= 0 || !results) return fail(SGDNET_ERR_ARG, "null specs/results or no fits");
  if (!basic_args(xa.n, xa.p, y, y_cols, why)) return fail(SGDNET_ERR_ARG, why);
  return guarded([&]() -> int {
    // row stride of `scores`: the caller sizes it from the controls it passed. A fit's resolved path never has more
    // than its own (or, with lambda_from, its source's) control.n_lambda values: FitPlan::build truncates a longer
    // given sequence to n_lambda.
    int max_lambda = 0;
    for (int i = 0; i < n_fits; ++i) max_lambda = std::max(max_lambda, specs[i].control.n_lambda);
    for (int i = 0; i < n_fits; ++i) {
      const sgdnet_fit_spec& s = specs[i];
      if (s.lambda_from >= i) return fail(SGDNET_ERR_ARG, "fit " + std::to_string(i) + ": lambda_from must name an earlier fit");
      std::string mwhy;
      if (scores && s.test_rows && !measure_ok(s.control.family, s.measure, mwhy)) return fail(SGDNET_ERR_ARG, "fit " + std::to_string(i) + ": " + mwhy);
      if (scores && s.test_rows && s.measure == SGDNET_MEASURE_AUC)
        return fail(SGDNET_ERR_ARG, "fit " + std::to_string(i) + ": type.measure 'auc' is scored one fit at a time (sgdnet_score_*): its tie-breaking "
                                    "draws come from the caller's generator on the calling thread");
      if (s.test_rows)
        for (int64_t q = 0; q < s.n_test; ++q)
          if (s.test_rows[q] < 0 || s.test_rows[q] >= xa.n) return fail(SGDNET_ERR_ARG, "test row id out of range");
    }
    // ONE engine for the whole batch: every fit is its own pipeline, whatever its kernel variant
    Engine eng;
    load_x(eng, xa, y, y_cols);
    // 1. designs (one per distinct row subset x standardize), serial
    for (int i = 0; i < n_fits; ++i) {
      sgdnet_fit_spec& s = specs[i];
      std::string err = eng.add_fit_design(s.train_rows, s.n_train, s.control, &s.rng, s.test_rows, s.n_test, s.measure);
      if (!err.empty()) return fail(SGDNET_ERR_ARG, "fit " + std::to_string(i) + ": " + err);
      eng.jobs.back().path_only = s.path_only != 0;
    }
    // 2. plans (response statistics, lambda path, step sizes) on the host cores, independent fits concurrently; a fit
    //    that takes its path from an earlier one (`lambda = lambda[[i]]`, R/cv_sgdnet.R:164, 186) goes in a later wave
    PhaseTimer pt;
    std::vector<std::string> errs(n_fits);
    std::vector<int> wave(n_fits, 0);
    int n_waves = 1;
    for (int i = 0; i < n_fits; ++i)
      if (specs[i].lambda_from >= 0) {
        wave[i] = wave[specs[i].lambda_from] + 1;
        n_waves = std::max(n_waves, wave[i] + 1);
      }
    for (int w = 0; w < n_waves; ++w) {
      std::vector<int> todo;
      for (int i = 0; i < n_fits; ++i)
        if (wave[i] == w) todo.push_back(i);
      std::atomic<size_t> next{0};
      auto work = [&]() {
        const bool on_device = cudaSetDevice(eng.dev_id) == cudaSuccess;
        for (;;) {
          const size_t q = next.fetch_add(1);
          if (q >= todo.size()) break;
          const int i = todo[q];
          if (!on_device) {
            errs[i] = "cudaSetDevice failed in a planning thread";
            continue;
          }
          sgdnet_control ctl = specs[i].control;
          if (specs[i].lambda_from >= 0) {
            const std::vector<double>& lam = eng.jobs[specs[i].lambda_from].plan.lambda;
            ctl.lambda = lam.data();
            ctl.lambda_len = static_cast<int32_t>(lam.size());
            ctl.n_lambda = ctl.lambda_len;
          }
          try {
            errs[i] = eng.build_plan(eng.jobs[i], specs[i].train_rows, ctl);
          } catch (const CudaFail& cf) {
            errs[i] = std::string(cf.what) + ": " + cudaGetErrorString(cf.e);
          } catch (const std::exception& e) {
            errs[i] = e.what();
          }
        }
      };
      const unsigned hc = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
      const size_t nt = std::min<size_t>(hc, todo.size());
      std::vector<std::thread> th;
      for (size_t k = 1; k < nt; ++k) th.emplace_back(work);
      work();
      for (auto& t : th) t.join();
      for (int i : todo)
        if (!errs[i].empty()) return fail(SGDNET_ERR_ARG, "fit " + std::to_string(i) + ": " + errs[i]);
    }
    pt.lap("plans (all fits)");
    // 3. device state
    size_t reserve = 0;
    for (int i = 0; i < n_fits; ++i)
      if (!eng.jobs[i].path_only) reserve += eng.fit_device_bytes(eng.jobs[i]);
    if (n_fits > 1) eng.arena.reserve(reserve);
    for (int i = 0; i < n_fits; ++i) {
      if (eng.jobs[i].path_only) eng.jobs[i].phase = Phase::Done;
      else eng.alloc_fit(eng.jobs[i]);
    }
    pt.lap("state alloc + upload (all fits)");
    if (scores) eng.upload_for_scoring();
    eng.finalize_batch();
    eng.run(-1, scores != nullptr);
    pt.lap("run (all fits)");
    for (int i = 0; i < n_fits; ++i) {
      if (!eng.jobs[i].path_only) eng.settle_rng(eng.jobs[i]);
      eng.fill_result(i, &results[i]);
      FitJob& j = eng.jobs[i];
      if (scores && j.scored)
        CK(cudaMemcpy(scores + size_t(i) * max_lambda, j.score_dev, sizeof(double) * std::min(j.plan.n_lambda, max_lambda),
                      cudaMemcpyDeviceToHost));
    }
    pt.lap("results to the host");
    return SGDNET_OK;
  });
}

int predict_or_score(const XArg& xa, const double* y, int32_t y_cols, int32_t family, const double* a0, const double* beta,
                     int32_t L, int32_t K, double* link, double* score, int32_t measure = SGDNET_MEASURE_DEVIANCE,
                     sgdnet_rng* rng = nullptr) {
  if (xa.n <= 0 || xa.p <= 0 || !a0 || !beta || L <= 0 || K <= 0 || K > 32) return fail(SGDNET_ERR_ARG, "bad predict arguments");
  if (score && (!y || y_cols <= 0)) return fail(SGDNET_ERR_ARG, "score needs y");
  std::string why;
  if (score && !measure_ok(family, measure, why)) return fail(SGDNET_ERR_ARG, why);
  if (score && measure == SGDNET_MEASURE_AUC) {
    // X * beta on the device; the rank statistic (a sort with tie-breaking draws from the caller's generator, which
    // must be called on the calling thread) on the host
    if (!rng || rng->kind == SGDNET_RNG_SEQUENCE || (rng->kind == SGDNET_RNG_CALLBACK && !rng->unif_rand))
      return fail(SGDNET_ERR_RNG, "type.measure 'auc' needs a generator for its tie-breaking draws (R/score.R:218)");
    std::vector<double> eta(static_cast<size_t>(L) * xa.n);
    const int rc = predict_or_score(xa, nullptr, 0, family, a0, beta, L, K, eta.data(), nullptr);
    if (rc != SGDNET_OK) return rc;
    for (int l = 0; l < L; ++l) score[l] = auc_one_lambda(eta.data() + static_cast<size_t>(l) * xa.n, y, xa.n, rng);
    return SGDNET_OK;
  }
  return guarded([&]() -> int {
    Engine eng;
    load_x(eng, xa, score ? y : nullptr, score ? y_cols : 1);
    std::vector<double> a0v(a0, a0 + size_t(L) * K), bv(beta, beta + size_t(L) * xa.p * K);
    double* a0d = eng.arena.upload(a0v);
    double* bd = eng.arena.upload(bv);
    double* link_dev = link ? eng.arena.alloc<double>(size_t(L) * K * xa.n, false) : nullptr;
    double* score_dev = score ? eng.arena.alloc<double>(L) : nullptr;
    eng.upload_raw();
    CK(cudaDeviceSynchronize());      // pageable uploads above (see finalize_batch)
    eng.predict_score(family, K, L, nullptr, xa.n, a0d, bd, score != nullptr, link_dev, score_dev, eng.stream, measure);
    CK(cudaStreamSynchronize(eng.stream));
    if (link) CK(cudaMemcpy(link, link_dev, sizeof(double) * size_t(L) * K * xa.n, cudaMemcpyDeviceToHost));
    if (score) CK(cudaMemcpy(score, score_dev, sizeof(double) * L, cudaMemcpyDeviceToHost));
    return SGDNET_OK;
  });
}

}  // namespace sgd

// ============================================================================================ extern "C"
using namespace sgd;

struct sgdnet_session {
  Engine eng;
};

extern "C" {

int sgdnet_abi_version(void) { return SGDNET_ABI_VERSION; }
const char* sgdnet_last_error(void) { return g_error.c_str(); }

void sgdnet_rng_set_seed(sgdnet_rng* rng, uint32_t seed) {
  std::memset(rng, 0, sizeof(*rng));
  rng->kind = SGDNET_RNG_MT;
  mt_seed(rng, seed);
}
double sgdnet_rng_unif(sgdnet_rng* rng) { return mt_unif(rng); }

int sgdnet_rng_indices(sgdnet_rng* rng, uint32_t n, int32_t n_epochs, uint32_t* seq, sgdnet_rng* states, int32_t on_host) {
  if (!rng || !seq || n == 0 || n_epochs <= 0) return fail(SGDNET_ERR_ARG, "bad rng_indices arguments");
  if (rng->kind != SGDNET_RNG_MT) return fail(SGDNET_ERR_ARG, "rng_indices needs an SGDNET_RNG_MT generator");
  return guarded([&]() -> int {
    MtState src{};
    std::memcpy(src.mt, rng->mt, sizeof(src.mt));
    src.mti = rng->mti;
    std::vector<MtState> snaps(size_t(n_epochs) + 1);
    const size_t total = size_t(n) * n_epochs;
    if (on_host) {
      mt_indices_host(&src, n, n_epochs, seq, snaps.data());
    } else {
      int count = 0;
      if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) throw CudaFail{cudaErrorNoDevice, "no CUDA device (no CPU fallback)"};
      Arena arena;
      MtState* src_d = arena.upload_from(&src, 1);
      MtState* snaps_d = arena.alloc<MtState>(snaps.size(), false);
      uint32_t* seq_d = arena.alloc<uint32_t>(total, false);
      CK(launch_mt_indices(src_d, n, n_epochs, seq_d, snaps_d, nullptr));
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(seq, seq_d, sizeof(uint32_t) * total, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(snaps.data(), snaps_d, sizeof(MtState) * snaps.size(), cudaMemcpyDeviceToHost));
    }
    for (int e = 0; e <= n_epochs; ++e) {
      sgdnet_rng* dst = (e == n_epochs) ? rng : nullptr;
      if (states) {
        states[e] = *rng;
        std::memcpy(states[e].mt, snaps[e].mt, sizeof(snaps[e].mt));
        states[e].mti = snaps[e].mti;
      }
      if (dst) {
        std::memcpy(dst->mt, snaps[e].mt, sizeof(snaps[e].mt));
        dst->mti = snaps[e].mti;
      }
    }
    return SGDNET_OK;
  });
}

void sgdnet_result_free(sgdnet_result* r) {
  if (!r) return;
  std::free(r->a0);
  std::free(r->beta);
  std::free(r->lambda);
  std::free(r->dev_ratio);
  std::free(r->return_codes);
  std::free(r->epochs);
  std::free(r->losses);
  std::free(r->losses_ptr);
  std::memset(r, 0, sizeof(*r));
}

int sgdnet_device_count(int* count) {
  if (!count) return fail(SGDNET_ERR_ARG, "null count");
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(SGDNET_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  }
  return SGDNET_OK;
}
int sgdnet_set_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(SGDNET_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
  return SGDNET_OK;
}

int sgdnet_fit_dense(const double* x, int64_t n, int64_t p, const double* y, int32_t y_cols,
                     const sgdnet_control* control, sgdnet_rng* rng, sgdnet_result* out) {
  if (!x) return fail(SGDNET_ERR_ARG, "x is null");
  return fit_single(XArg{false, x, nullptr, nullptr, nullptr, n, p}, y, y_cols, control, rng, out);
}
int sgdnet_fit_sparse(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x, int64_t n, int64_t p,
                      const double* y, int32_t y_cols, const sgdnet_control* control, sgdnet_rng* rng,
                      sgdnet_result* out) {
  if (!csc_i || !csc_p || !csc_x) return fail(SGDNET_ERR_ARG, "sparse x is null");
  return fit_single(XArg{true, nullptr, csc_i, csc_p, csc_x, n, p}, y, y_cols, control, rng, out);
}

int sgdnet_fit_batch_dense(const double* x, int64_t n, int64_t p, const double* y, int32_t y_cols,
                           sgdnet_fit_spec* specs, int32_t n_fits, sgdnet_result* results, double* scores) {
  if (!x) return fail(SGDNET_ERR_ARG, "x is null");
  return fit_batch(XArg{false, x, nullptr, nullptr, nullptr, n, p}, y, y_cols, specs, n_fits, results, scores);
}
int sgdnet_fit_batch_sparse(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x, int64_t n, int64_t p,
                            const double* y, int32_t y_cols, sgdnet_fit_spec* specs, int32_t n_fits,
                            sgdnet_result* results, double* scores) {
  if (!csc_i || !csc_p || !csc_x) return fail(SGDNET_ERR_ARG, "sparse x is null");
  return fit_batch(XArg{true, nullptr, csc_i, csc_p, csc_x, n, p}, y, y_cols, specs, n_fits, results, scores);
}

int sgdnet_predict_dense(const double* x, int64_t n, int64_t p, const double* a0, const double* beta, int32_t n_lambda,
                         int32_t n_classes, double* link) {
  if (!x || !link) return fail(SGDNET_ERR_ARG, "null x or link");
  return predict_or_score(XArg{false, x, nullptr, nullptr, nullptr, n, p}, nullptr, 0, 0, a0, beta, n_lambda, n_classes, link, nullptr);
}
int sgdnet_predict_sparse(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x, int64_t n, int64_t p,
                          const double* a0, const double* beta, int32_t n_lambda, int32_t n_classes, double* link) {
  if (!csc_i || !csc_p || !csc_x || !link) return fail(SGDNET_ERR_ARG, "null x or link");
  return predict_or_score(XArg{true, nullptr, csc_i, csc_p, csc_x, n, p}, nullptr, 0, 0, a0, beta, n_lambda, n_classes, link, nullptr);
}
int sgdnet_score_deviance_dense(const double* x, int64_t n, int64_t p, const double* y, int32_t y_cols, int32_t family,
                                const double* a0, const double* beta, int32_t n_lambda, int32_t n_classes, double* score) {
  if (!x || !score) return fail(SGDNET_ERR_ARG, "null x or score");
  return predict_or_score(XArg{false, x, nullptr, nullptr, nullptr, n, p}, y, y_cols, family, a0, beta, n_lambda, n_classes, nullptr, score);
}
int sgdnet_score_deviance_sparse(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x, int64_t n, int64_t p,
                                 const double* y, int32_t y_cols, int32_t family, const double* a0, const double* beta,
                                 int32_t n_lambda, int32_t n_classes, double* score) {
  if (!csc_i || !csc_p || !csc_x || !score) return fail(SGDNET_ERR_ARG, "null x or score");
  return predict_or_score(XArg{true, nullptr, csc_i, csc_p, csc_x, n, p}, y, y_cols, family, a0, beta, n_lambda, n_classes, nullptr, score);
}

int sgdnet_score_dense(const double* x, int64_t n, int64_t p, const double* y, int32_t y_cols, int32_t family, int32_t measure,
                       const double* a0, const double* beta, int32_t n_lambda, int32_t n_classes, sgdnet_rng* rng, double* score) {
  if (!x || !score) return fail(SGDNET_ERR_ARG, "null x or score");
  return predict_or_score(XArg{false, x, nullptr, nullptr, nullptr, n, p}, y, y_cols, family, a0, beta, n_lambda, n_classes, nullptr, score, measure, rng);
}
int sgdnet_score_sparse(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x, int64_t n, int64_t p, const double* y,
                        int32_t y_cols, int32_t family, int32_t measure, const double* a0, const double* beta, int32_t n_lambda,
                        int32_t n_classes, sgdnet_rng* rng, double* score) {
  if (!csc_i || !csc_p || !csc_x || !score) return fail(SGDNET_ERR_ARG, "null x or score");
  return predict_or_score(XArg{true, nullptr, csc_i, csc_p, csc_x, n, p}, y, y_cols, family, a0, beta, n_lambda, n_classes, nullptr, score, measure, rng);
}

// ---- stepping interface
static int session_create(const XArg& xa, const double* y, int32_t y_cols, const sgdnet_control* control,
                          sgdnet_session** out) {
  std::string why;
  if (!control || !out) return fail(SGDNET_ERR_ARG, "null control or session pointer");
  if (!basic_args(xa.n, xa.p, y, y_cols, why)) return fail(SGDNET_ERR_ARG, why);
  *out = nullptr;
  return guarded([&]() -> int {
    std::unique_ptr<sgdnet_session> s(new sgdnet_session());
    load_x(s->eng, xa, y, y_cols);
    static thread_local sgdnet_rng placeholder;   // replaced per call
    placeholder.kind = SGDNET_RNG_MT;
    std::string err = s->eng.add_fit(nullptr, xa.n, *control, &placeholder, nullptr, 0);
    if (!err.empty()) return fail(SGDNET_ERR_ARG, err);
    s->eng.finalize_batch();
    *out = s.release();
    return SGDNET_OK;
  });
}

int sgdnet_session_create_dense(const double* x, int64_t n, int64_t p, const double* y, int32_t y_cols,
                                const sgdnet_control* control, sgdnet_session** out) {
  if (!x) return fail(SGDNET_ERR_ARG, "x is null");
  return session_create(XArg{false, x, nullptr, nullptr, nullptr, n, p}, y, y_cols, control, out);
}
int sgdnet_session_create_sparse(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x, int64_t n, int64_t p,
                                 const double* y, int32_t y_cols, const sgdnet_control* control, sgdnet_session** out) {
  if (!csc_i || !csc_p || !csc_x) return fail(SGDNET_ERR_ARG, "sparse x is null");
  return session_create(XArg{true, nullptr, csc_i, csc_p, csc_x, n, p}, y, y_cols, control, out);
}

int sgdnet_session_run_epochs(sgdnet_session* s, int32_t lambda_ind, int32_t n_epochs, sgdnet_rng* rng, float* device_ms) {
  if (!s || !rng || n_epochs <= 0) return fail(SGDNET_ERR_ARG, "bad session arguments");
  return guarded([&]() -> int {
    Engine& e = s->eng;
    FitJob& j = e.jobs[0];
    if (lambda_ind < 0 || lambda_ind >= j.plan.n_lambda) return fail(SGDNET_ERR_ARG, "lambda index out of range");
    j.rng = rng;
    if (j.device_rng != (rng->kind == SGDNET_RNG_MT && !e.host_rng)) return fail(SGDNET_ERR_ARG, "session created for another generator kind");
    e.upload_rng(j);
    // measurement mode: run exactly n_epochs at lambda_ind, never leave kRunning
    const double solver_before = j.seconds_solver;
    int left = n_epochs;
    while (left > 0) {
      const int ne = std::min(left, j.epochs_per_launch);
      j.mirror->lambda_ind = lambda_ind;
      j.mirror->status = kRunning;
      j.mirror->it_outer = (j.mirror->it_outer == 0) ? 0u : 1u;   // keep "new lambda" only for the first call
      CK(cudaMemcpyAsync(j.prog_ptr, j.mirror, sizeof(Progress), cudaMemcpyHostToDevice, j.st));
      e.submit_solver(j, 1, ne);
      e.wait_round(j);
      e.solver_done(j);
      j.needs_finish = false;
      j.phase = Phase::Idle;
      j.mirror->it_outer = 1;
      left -= ne;
    }
    e.settle_rng(j);
    if (device_ms) *device_ms = static_cast<float>((j.seconds_solver - solver_before) * 1e3);
    return SGDNET_OK;
  });
}

int sgdnet_session_fit_lambda(sgdnet_session* s, int32_t lambda_ind, sgdnet_rng* rng, uint32_t* epochs, int32_t* converged) {
  if (!s || !rng) return fail(SGDNET_ERR_ARG, "bad session arguments");
  return guarded([&]() -> int {
    Engine& e = s->eng;
    FitJob& j = e.jobs[0];
    if (lambda_ind != j.mirror->lambda_ind) return fail(SGDNET_ERR_ARG, "lambdas must be fitted in path order");
    j.rng = rng;
    if (j.device_rng != (rng->kind == SGDNET_RNG_MT && !e.host_rng)) return fail(SGDNET_ERR_ARG, "session created for another generator kind");
    e.run(lambda_ind);
    e.settle_rng(j);
    uint32_t ep = 0, code = 0;
    CK(cudaMemcpy(&ep, j.dev.epochs + lambda_ind, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&code, j.dev.codes + lambda_ind, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (epochs) *epochs = ep;
    if (converged) *converged = code == 0;
    return SGDNET_OK;
  });
}

int sgdnet_session_finish_lambda(sgdnet_session* s, int32_t lambda_ind, float* device_ms) {
  if (!s) return fail(SGDNET_ERR_ARG, "null session");
  return guarded([&]() -> int {
    Engine& e = s->eng;
    FitJob& j = e.jobs[0];
    // measurement entry: force the deviance + rescale pass for lambda_ind on the current state
    j.mirror->lambda_ind = lambda_ind;
    j.mirror->status = kLambdaDone;
    CK(cudaMemcpyAsync(j.prog_ptr, j.mirror, sizeof(Progress), cudaMemcpyHostToDevice, j.st));
    const double before = j.seconds_dev;
    e.submit_finish(j);
    e.wait_round(j);
    e.finish_done(j);
    j.phase = Phase::Idle;
    if (device_ms) *device_ms = static_cast<float>((j.seconds_dev - before) * 1e3);
    return SGDNET_OK;
  });
}

int sgdnet_session_result(sgdnet_session* s, sgdnet_result* out) {
  if (!s || !out) return fail(SGDNET_ERR_ARG, "null session or result");
  return guarded([&]() -> int {
    s->eng.fill_result(0, out);
    return SGDNET_OK;
  });
}

void sgdnet_session_destroy(sgdnet_session* s) { delete s; }

}  // extern "C"
