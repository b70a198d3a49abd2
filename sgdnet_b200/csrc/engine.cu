// engine.cu — the extern "C" boundary (include/sgdnet_b200.h) and the lambda-path driver behind it.
//
// What replaces what: sgdnet_fit_dense / sgdnet_fit_sparse are SgdnetDense / SgdnetSparse (reference
// src/sgdnet.cpp:359-375); the driver below is SetupSgdnet's lambda loop (src/sgdnet.cpp:217-273) turned into a
// per-fit state machine that lives on the device (Progress): every "round" the host launches
//     lag-scaling table (new lambda only) -> SAGA epochs -> [debug epoch loss] -> deviance + rescale + archive
// once for ALL fits of a batch (one CTA per fit for the solver, grid-wide passes for the streaming kernels), then
// reads the few bytes of Progress back. Warm-start state never leaves HBM (src/sgdnet.cpp:186-198).
// The sampling sequence is produced on the host in the reference's order (floor(runif(0,n)) per update) and
// uploaded ahead of the launch; draws that a converged epoch did not consume are kept for the next lambda, and the
// generator handed back to the caller has advanced by exactly n * npasses draws.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sgdnet_b200.h"
#include "host_setup.h"
#include "kernels.h"

namespace sgd {

thread_local std::string g_error;

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
};

struct FitJob {
  std::shared_ptr<HostDesign> design;
  DeviceDesign ddev;
  FitPlan plan;
  sgdnet_rng* rng = nullptr;
  FitDev dev{};                       // host mirror
  // index stream
  std::vector<uint32_t> pending;      // generated, not yet consumed
  size_t pending_head = 0;
  uint64_t consumed = 0;              // draws consumed by finished epochs
  std::vector<std::pair<uint64_t, sgdnet_rng>> marks;   // (draws generated before, generator state) per block
  uint64_t generated = 0;
  uint32_t* seq_dev = nullptr;
  uint32_t* seq_pin = nullptr;
  uint64_t* dep_dev = nullptr;        // sparse K == 1: conflict codes of the staged sequence (wave_deps_kernel)
  uint8_t* dup_dev = nullptr;
  int epochs_per_launch = 1;
  bool done = false;
  // scoring of held-out rows after the fit (cv)
  const int32_t* test_rows = nullptr;
  int64_t n_test = 0;
};

enum class Variant { Dense, SparseK1, SparseGeneric };

struct Engine {
  Arena arena;
  RawX raw;
  std::vector<double> y_cm;           // caller's y, n x Ky column-major
  int Ky = 1;
  std::map<std::pair<const int32_t*, int>, std::pair<std::shared_ptr<HostDesign>, DeviceDesign>> designs;
  std::vector<FitJob> jobs;
  FitDev* fits_dev = nullptr;
  Progress* prog_dev = nullptr;
  Progress* prog_host = nullptr;      // pinned
  RoundArgs* args_dev = nullptr;
  RoundArgs* args_host = nullptr;     // pinned
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
  int loss_blocks = 1;
  int sms = 148;
  int64_t max_rows_per_launch = 1;
  bool trace_rounds = std::getenv("SGDNET_TRACE_ROUNDS") != nullptr;   // per-round device times on stderr
  size_t dense_smem = 0;
  unsigned dense_kts = 0, dense_pens = 0;   // class-count buckets / penalties present (dense kernel instantiations)
  Variant variant = Variant::Dense;
  bool any_debug = false;
  uint64_t launches = 0;
  double seconds_solver = 0.0, seconds_dev = 0.0, seconds_setup = 0.0;
  double t_begin = 0.0;
  // raw design for scoring (device)
  DeviceDesign raw_dev;
  bool raw_uploaded = false;
  double* yraw_dev = nullptr;

  Engine() {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) throw CudaFail{e == cudaSuccess ? cudaErrorNoDevice : e, "no CUDA device (no CPU fallback)"};
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    arena.stream = stream;
    CK(cudaEventCreate(&ev0));
    CK(cudaEventCreate(&ev1));
    CK(cudaEventCreate(&ev2));
    t_begin = now_s();
  }
  ~Engine() {
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (ev2) cudaEventDestroy(ev2);
    if (stream) cudaStreamDestroy(stream);
  }

  // ---------------------------------------------------------------------------------- designs
  std::pair<std::shared_ptr<HostDesign>, DeviceDesign> get_design(const int32_t* rows, int64_t n_rows, bool standardize) {
    auto key = std::make_pair(rows, standardize ? 1 : 0);
    auto it = designs.find(key);
    if (it != designs.end()) return it->second;
    auto hd = std::make_shared<HostDesign>();
    PhaseTimer pt;
    hd->build(raw, rows, n_rows, standardize);
    pt.lap("design build (host)");
    DeviceDesign dd;
    if (hd->sparse) {
      dd.rows = arena.upload_from(hd->rows_v, static_cast<size_t>(hd->n));
      dd.ci = arena.upload_from(hd->ci_v, hd->n_entries);
      dd.cv = arena.upload_from(hd->cv_v, hd->n_entries);
    } else {
      dd.xd = arena.upload(hd->xd);
    }
    pt.lap("design upload");
    dd.c = arena.upload(hd->c);
    dd.x_center = arena.upload(hd->x_center);
    dd.x_scale = arena.upload(hd->x_scale);
    designs[key] = {hd, dd};
    return {hd, dd};
  }

  // ---------------------------------------------------------------------------------- one fit
  std::string add_fit(const int32_t* rows, int64_t n_rows, const sgdnet_control& ctl, sgdnet_rng* rng,
                      const int32_t* test_rows, int64_t n_test) {
    if (ctl.n_lambda <= 0) return "n_lambda must be positive";
    if (!rng) return "rng is null";
    if (rows)
      for (int64_t i = 0; i < n_rows; ++i)
        if (rows[i] < 0 || rows[i] >= raw.n) return "train row id out of range";
    FitJob job;
    auto dsn = get_design(rows, n_rows, ctl.standardize != 0);
    job.design = dsn.first;
    job.ddev = dsn.second;
    const HostDesign& d = *job.design;
    // response restricted to the fit's rows
    std::vector<double> ysub(static_cast<size_t>(d.n) * Ky);
    for (int k = 0; k < Ky; ++k)
      for (int64_t i = 0; i < d.n; ++i)
        ysub[static_cast<size_t>(k) * d.n + i] = y_cm[static_cast<size_t>(k) * raw.n + (rows ? rows[i] : i)];
    PhaseTimer pt;
    std::string err = job.plan.build(d, std::move(ysub), Ky, ctl);
    if (!err.empty()) return err;
    pt.lap("plan (lambda path, steps)");
    job.rng = rng;
    job.test_rows = test_rows;
    job.n_test = n_test;

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
    f.st = (d.sparse && K == 1 && !f.standardize) ? arena.alloc<FeatState>(p) : nullptr;
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
    any_debug = any_debug || pl.debug;

    // epochs per launch: amortise the round trip for small problems; the callback generator and the debug loss need
    // a host visit after every epoch
    int epl = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(16, 200000 / std::max<int64_t>(1, d.n))));
    if (const char* env = std::getenv("SGDNET_EPOCHS_PER_LAUNCH")) epl = std::max(1, std::atoi(env));   // tuning knob
    if (pl.debug || rng->kind == SGDNET_RNG_CALLBACK) epl = 1;
    epl = static_cast<int>(std::min<uint32_t>(static_cast<uint32_t>(epl), std::max<uint32_t>(1u, pl.max_iter)));
    job.epochs_per_launch = epl;
    job.seq_dev = arena.alloc<uint32_t>(size_t(epl) * d.n, false);
    job.seq_pin = arena.host<uint32_t>(size_t(epl) * d.n);
    if (d.sparse && K == 1 && !f.standardize) {
      job.dep_dev = arena.alloc<uint64_t>(size_t(epl) * d.n * 32, false);
      job.dup_dev = arena.alloc<uint8_t>(size_t(epl) * d.n, false);
    }
    jobs.push_back(std::move(job));
    pt.lap("state alloc + upload");
    return "";
  }

  void finalize_batch() {
    const int nf = static_cast<int>(jobs.size());
    // one kernel variant per batch
    const FitJob& j0 = jobs[0];
    if (!j0.dev.sparse) variant = Variant::Dense;
    else variant = (j0.dev.K == 1 && !j0.dev.standardize) ? Variant::SparseK1 : Variant::SparseGeneric;
    int64_t max_n = 0;
    for (auto& j : jobs) max_n = std::max(max_n, j.dev.n);
    if (variant == Variant::Dense) {
      // the launch's dynamic shared memory: what the most demanding fit of the batch needs (each CTA decides from
      // its own K and p whether its state fits it)
      dense_smem = 0;
      for (auto& j : jobs) {
        int in_smem = 0;
        const size_t need = dense_smem_bytes(j.dev.K, j.dev.p, j.dev.ld, &in_smem);
        if (need > dense_smem_budget())
          throw std::invalid_argument("dense x with p = " + std::to_string(j.dev.p) + " columns: a row ring of " +
                                      std::to_string(need) + " bytes exceeds one SM's shared memory (" +
                                      std::to_string(dense_smem_budget()) + "); this build handles dense p up to about 7000");
        dense_smem = std::max(dense_smem, need);
      }
      for (auto& j : jobs) {
        dense_kts |= static_cast<unsigned>(dense_kt_bucket(j.dev.K));
        dense_pens |= 1u << j.dev.penalty;
      }
    }
    // streaming passes: enough CTAs to fill the GPU across the fits of the batch, at least one per fit
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
    for (auto& j : jobs) max_rows_per_launch = std::max<int64_t>(max_rows_per_launch, j.dev.n * j.epochs_per_launch);
    const int64_t want = std::max<int64_t>(1, (int64_t(sms) * 4 + nf - 1) / nf);
    loss_blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(want, (max_n + 7) / 8)));
    for (auto& j : jobs) j.dev.partials = arena.alloc<double>(loss_blocks);

    std::vector<FitDev> mirror(nf);
    for (int i = 0; i < nf; ++i) mirror[i] = jobs[i].dev;
    fits_dev = arena.upload(mirror);
    prog_dev = arena.alloc<Progress>(nf, false);
    prog_host = arena.host<Progress>(nf);
    std::memset(prog_host, 0, sizeof(Progress) * nf);
    for (int i = 0; i < nf; ++i) prog_host[i].wscale = 1.0;
    CK(cudaMemcpyAsync(prog_dev, prog_host, sizeof(Progress) * nf, cudaMemcpyHostToDevice, stream));
    args_dev = arena.alloc<RoundArgs>(nf);
    args_host = arena.host<RoundArgs>(nf);
    seconds_setup = now_s() - t_begin;
  }

  // ---------------------------------------------------------------------------------- index stream
  // Makes sure `need` undrawn-by-the-device indices are pending (host memory). May run on a helper thread while the
  // device works: it touches only this job's generator and buffers.
  bool ensure_pending(FitJob& j, size_t need) {
    size_t have = j.pending.size() - j.pending_head;
    if (have >= need) return true;
    if (j.pending_head > 0) {
      j.pending.erase(j.pending.begin(), j.pending.begin() + j.pending_head);
      j.pending_head = 0;
    }
    const size_t add = need - have;
    // keep every mark from the newest one that is not ahead of the consumed position: settle_rng restarts from it
    // (`consumed` only moves between rounds, after the helper threads have been joined)
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

  // While the device runs a round: draw the next round's indices for every fit that may need them (a fit consumes
  // at most one launch's worth per round), spread over host threads. The callback generator is never drawn ahead
  // (it must be called on the caller's thread, exactly as often as the reference would call it).
  void prefetch_indices() {
    std::vector<FitJob*> todo;
    for (size_t i = 0; i < jobs.size(); ++i) {
      FitJob& j = jobs[i];
      if (j.done || args_host[i].n_epochs == 0 || j.rng->kind == SGDNET_RNG_CALLBACK) continue;
      if (j.rng->kind == SGDNET_RNG_SEQUENCE &&
          j.rng->seq_len - j.rng->seq_pos < static_cast<int64_t>(2 * size_t(j.epochs_per_launch) * j.dev.n)) continue;
      todo.push_back(&j);
    }
    if (todo.empty()) return;
    const unsigned hc = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const int K = static_cast<int>(std::min<size_t>(hc, todo.size()));
    auto work = [&](int k) {
      for (size_t q = k; q < todo.size(); q += K) {
        FitJob& j = *todo[q];
        (void)ensure_pending(j, 2 * size_t(j.epochs_per_launch) * j.dev.n);
      }
    };
    if (K == 1) {
      work(0);
      return;
    }
    std::vector<std::thread> th;
    for (int k = 1; k < K; ++k) th.emplace_back(work, k);
    work(0);
    for (auto& t : th) t.join();
  }

  // Give the caller's generator back advanced by exactly the draws the fit consumed.
  void settle_rng(FitJob& j) {
    if (j.generated == j.consumed) return;
    if (j.rng->kind == SGDNET_RNG_SEQUENCE) {
      j.rng->seq_pos -= static_cast<int64_t>(j.generated - j.consumed);
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
        return;
      }
    }
    throw std::runtime_error("internal: no generator mark at or before the consumed position");
  }

  // ---------------------------------------------------------------------------------- rounds
  // Runs every job to the end of its path (or, with `only_lambda` >= 0, until that lambda is finished).
  void run(int only_lambda = -1) {
    const int nf = static_cast<int>(jobs.size());
    for (;;) {
      int active = 0;
      for (int i = 0; i < nf; ++i) {
        FitJob& j = jobs[i];
        const Progress& pg = prog_host[i];
        const bool stop_here = (only_lambda >= 0 && pg.lambda_ind > only_lambda);
        if (j.done || pg.status == kFitDone || stop_here) {
          args_host[i] = RoundArgs{nullptr, nullptr, nullptr, 0, 0};
          continue;
        }
        const uint32_t left = j.plan.max_iter - pg.it_outer;
        const int ne = static_cast<int>(std::min<uint32_t>(static_cast<uint32_t>(j.epochs_per_launch), std::max<uint32_t>(left, 1u)));
        if (!stage_indices(j, ne)) {
          g_error = "sampling-index source exhausted";
          throw CudaFail{cudaSuccess, "rng"};
        }
        CK(cudaMemcpyAsync(j.seq_dev, j.seq_pin, size_t(ne) * j.dev.n * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
        args_host[i] = RoundArgs{j.seq_dev, j.dep_dev, j.dup_dev, ne, 0};
        ++active;
      }
      if (active == 0) break;
      CK(cudaMemcpyAsync(args_dev, args_host, sizeof(RoundArgs) * nf, cudaMemcpyHostToDevice, stream));
      if (variant != Variant::Dense) {
        CK(launch_lag_scaling(nf, fits_dev, prog_dev, stream));
        ++launches;
      }
      CK(cudaEventRecord(ev0, stream));
      if (variant == Variant::SparseK1) {
        CK(launch_wave_deps(nf, fits_dev, prog_dev, args_dev, max_rows_per_launch, sms, stream));
        ++launches;
      }
      if (variant == Variant::Dense)
        CK(launch_saga_dense(nf, dense_kts, dense_pens, dense_smem, fits_dev, prog_dev, args_dev, stream));
      else
        CK(launch_saga_sparse(nf, variant == Variant::SparseK1, fits_dev, prog_dev, args_dev, stream));
      ++launches;
      CK(cudaEventRecord(ev1, stream));
      if (any_debug) {
        CK(launch_epoch_loss(nf, fits_dev, prog_dev, args_dev, loss_blocks, stream));
        launches += 2;
      }
      CK(launch_finish_lambda(nf, fits_dev, prog_dev, loss_blocks, stream));
      launches += 2;
      CK(cudaEventRecord(ev2, stream));
      CK(cudaMemcpyAsync(prog_host, prog_dev, sizeof(Progress) * nf, cudaMemcpyDeviceToHost, stream));
      prefetch_indices();
      CK(cudaStreamSynchronize(stream));
      float ms_solver = 0.f, ms_dev = 0.f;
      CK(cudaEventElapsedTime(&ms_solver, ev0, ev1));
      CK(cudaEventElapsedTime(&ms_dev, ev1, ev2));
      seconds_solver += ms_solver * 1e-3;
      seconds_dev += ms_dev * 1e-3;
      if (trace_rounds) std::fprintf(stderr, "[sgdnet_b200] round: solver %.3f ms, passes %.3f ms, fits active %d\n", ms_solver, ms_dev, active);
      for (int i = 0; i < nf; ++i) {
        if (args_host[i].n_epochs == 0) continue;
        FitJob& j = jobs[i];
        const uint64_t used = uint64_t(prog_host[i].epochs_last_launch) * uint64_t(j.dev.n);
        j.pending_head += used;
        j.consumed += used;
        if (prog_host[i].status == kFitDone) j.done = true;
      }
    }
  }

  // ---------------------------------------------------------------------------------- results
  void fill_result(int i, sgdnet_result* out) {
    FitJob& j = jobs[i];
    const FitPlan& pl = j.plan;
    const int L = pl.n_lambda, K = pl.K, p = j.dev.p;
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
    CK(cudaMemcpy(out->a0, j.dev.a0_arch, sizeof(double) * L * K, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out->beta, j.dev.beta_arch, sizeof(double) * size_t(L) * p * K, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out->dev_ratio, j.dev.dev_ratio, sizeof(double) * L, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out->return_codes, j.dev.codes, sizeof(uint32_t) * L, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out->epochs, j.dev.epochs, sizeof(uint32_t) * L, cudaMemcpyDeviceToHost));
    std::memcpy(out->lambda, pl.lambda.data(), sizeof(double) * L);
    out->nulldev = pl.nulldev;
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
    out->seconds_solver = seconds_solver;
    out->seconds_deviance = seconds_dev;
    out->seconds_setup = seconds_setup;
    out->seconds_total = now_s() - t_begin;
    out->kernel_launches = launches;
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

  // score/link of rows `row_ids` (device pointer or null) under coefficients (a0_dev, beta_dev)
  void predict_score(int family, int K, int L, const int32_t* row_ids_dev, int64_t n_rows, const double* a0_dev,
                     const double* beta_dev, bool with_y, double* link_dev, double* score_dev) {
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
    int sms = 148, dev_id = 0;
    cudaGetDevice(&dev_id);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
    const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(int64_t(sms) * 4, (n_rows + 7) / 8)));
    a.partials = arena.alloc<double>(size_t(blocks) * L);
    a.score = score_dev;
    double* bt = arena.alloc<double>(size_t(hd.p) * L * K, false);
    CK(launch_predict_score(a, bt, blocks, stream));
    launches += score_dev ? 3 : 2;
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
  if (xa.sparse) eng.raw.from_csc(xa.ci, xa.cp, xa.cx, xa.n, xa.p);
  else eng.raw.from_dense(xa.x, xa.n, xa.p);
  pt.lap("CSC -> CSR");
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

int fit_batch(const XArg& xa, const double* y, int32_t y_cols, sgdnet_fit_spec* specs, int32_t n_fits,
              sgdnet_result* results, double* scores) {
  std::string why;
  if (!specs || n_fits <= 0 || !results) return fail(SGDNET_ERR_ARG, "null specs/results or no fits");
  if (!basic_args(xa.n, xa.p, y, y_cols, why)) return fail(SGDNET_ERR_ARG, why);
  return guarded([&]() -> int {
    // group fits by kernel variant; each group is one batch of concurrent CTAs
    std::vector<int> order(n_fits);
    for (int i = 0; i < n_fits; ++i) order[i] = i;
    auto variant_of = [&](const sgdnet_fit_spec& s) {
      if (!xa.sparse) return s.control.n_classes == 1 ? 0 : 1;
      return (s.control.n_classes == 1 && !s.control.standardize) ? 2 : 3;
    };
    // row stride of `scores`: the caller sizes it from the controls it passed. A fit's resolved path never has more
    // than its own (or, with lambda_from, its source's) control.n_lambda values: FitPlan::build truncates a longer
    // given sequence to n_lambda.
    int max_lambda = 0;
    for (int i = 0; i < n_fits; ++i) max_lambda = std::max(max_lambda, specs[i].control.n_lambda);
    for (int v = 0; v < 4; ++v) {
      std::vector<int> group;
      for (int i = 0; i < n_fits; ++i)
        if (variant_of(specs[i]) == v) group.push_back(i);
      if (group.empty()) continue;
      Engine eng;
      load_x(eng, xa, y, y_cols);
      std::map<int, int> job_of;   // spec index -> job index inside this engine
      for (int i : group) {
        sgdnet_fit_spec& s = specs[i];
        sgdnet_control ctl = s.control;
        if (s.lambda_from >= 0) {
          // lambda = the path of an earlier fit of the batch (R/cv_sgdnet.R:164, 186), known after that fit's setup
          auto src = job_of.find(s.lambda_from);
          if (s.lambda_from >= i || src == job_of.end())
            return fail(SGDNET_ERR_ARG, "fit " + std::to_string(i) + ": lambda_from must name an earlier fit of the same kind");
          const std::vector<double>& lam = eng.jobs[src->second].plan.lambda;
          ctl.lambda = lam.data();
          ctl.lambda_len = static_cast<int32_t>(lam.size());
          ctl.n_lambda = ctl.lambda_len;
        }
        std::string err = eng.add_fit(s.train_rows, s.n_train, ctl, &s.rng, s.test_rows, s.n_test);
        if (!err.empty()) return fail(SGDNET_ERR_ARG, "fit " + std::to_string(i) + ": " + err);
        job_of[i] = static_cast<int>(eng.jobs.size()) - 1;
      }
      eng.finalize_batch();
      eng.run();
      for (size_t g = 0; g < group.size(); ++g) {
        eng.settle_rng(eng.jobs[g]);
        eng.fill_result(static_cast<int>(g), &results[group[g]]);
      }
      if (scores) {
        // score(fit, x_test, y_test, "deviance") for every fit with held-out rows (R/cv_sgdnet.R:197-198)
        std::map<const int32_t*, int32_t*> test_dev;
        for (size_t g = 0; g < group.size(); ++g) {
          FitJob& j = eng.jobs[g];
          if (!j.test_rows || j.n_test <= 0) continue;
          for (int64_t q = 0; q < j.n_test; ++q)
            if (j.test_rows[q] < 0 || j.test_rows[q] >= xa.n) return fail(SGDNET_ERR_ARG, "test row id out of range");
          int32_t*& td = test_dev[j.test_rows];
          if (!td) {
            td = eng.arena.alloc<int32_t>(j.n_test, false);
            CK(cudaMemcpy(td, j.test_rows, sizeof(int32_t) * j.n_test, cudaMemcpyHostToDevice));
          }
          const int L = j.plan.n_lambda;
          double* score_dev = eng.arena.alloc<double>(L);
          eng.predict_score(j.plan.family, j.plan.K, L, td, j.n_test, j.dev.a0_arch, j.dev.beta_arch, true, nullptr, score_dev);
          CK(cudaStreamSynchronize(eng.stream));
          CK(cudaMemcpy(scores + size_t(group[g]) * max_lambda, score_dev, sizeof(double) * std::min(L, max_lambda), cudaMemcpyDeviceToHost));
        }
      }
    }
    return SGDNET_OK;
  });
}

int predict_or_score(const XArg& xa, const double* y, int32_t y_cols, int32_t family, const double* a0, const double* beta,
                     int32_t L, int32_t K, double* link, double* score) {
  if (xa.n <= 0 || xa.p <= 0 || !a0 || !beta || L <= 0 || K <= 0 || K > 32) return fail(SGDNET_ERR_ARG, "bad predict arguments");
  if (score && (!y || y_cols <= 0)) return fail(SGDNET_ERR_ARG, "score needs y");
  return guarded([&]() -> int {
    Engine eng;
    load_x(eng, xa, score ? y : nullptr, score ? y_cols : 1);
    std::vector<double> a0v(a0, a0 + size_t(L) * K), bv(beta, beta + size_t(L) * xa.p * K);
    double* a0d = eng.arena.upload(a0v);
    double* bd = eng.arena.upload(bv);
    double* link_dev = link ? eng.arena.alloc<double>(size_t(L) * K * xa.n, false) : nullptr;
    double* score_dev = score ? eng.arena.alloc<double>(L) : nullptr;
    eng.predict_score(family, K, L, nullptr, xa.n, a0d, bd, score != nullptr, link_dev, score_dev);
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
    s->eng.raw.release_rows();
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
    // measurement mode: run exactly n_epochs at lambda_ind, never leave kRunning
    float total_ms = 0.f;
    int left = n_epochs;
    while (left > 0) {
      const int ne = std::min(left, j.epochs_per_launch);
      e.prog_host[0].lambda_ind = lambda_ind;
      e.prog_host[0].status = kRunning;
      e.prog_host[0].it_outer = (e.prog_host[0].it_outer == 0) ? 0u : 1u;   // keep "new lambda" only for the first call
      CK(cudaMemcpyAsync(e.prog_dev, e.prog_host, sizeof(Progress), cudaMemcpyHostToDevice, e.stream));
      if (!e.stage_indices(j, ne)) return fail(SGDNET_ERR_RNG, "sampling-index source exhausted");
      CK(cudaMemcpyAsync(j.seq_dev, j.seq_pin, size_t(ne) * j.dev.n * sizeof(uint32_t), cudaMemcpyHostToDevice, e.stream));
      e.args_host[0] = RoundArgs{j.seq_dev, j.dep_dev, j.dup_dev, ne, 1};
      CK(cudaMemcpyAsync(e.args_dev, e.args_host, sizeof(RoundArgs), cudaMemcpyHostToDevice, e.stream));
      if (e.variant != Variant::Dense) { CK(launch_lag_scaling(1, e.fits_dev, e.prog_dev, e.stream)); ++e.launches; }
      CK(cudaEventRecord(e.ev0, e.stream));
      if (e.variant == Variant::SparseK1) {
        CK(launch_wave_deps(1, e.fits_dev, e.prog_dev, e.args_dev, e.max_rows_per_launch, e.sms, e.stream));
        ++e.launches;
      }
      if (e.variant == Variant::Dense)
        CK(launch_saga_dense(1, e.dense_kts, e.dense_pens, e.dense_smem, e.fits_dev, e.prog_dev, e.args_dev, e.stream));
      else
        CK(launch_saga_sparse(1, e.variant == Variant::SparseK1, e.fits_dev, e.prog_dev, e.args_dev, e.stream));
      ++e.launches;
      CK(cudaEventRecord(e.ev1, e.stream));
      CK(cudaMemcpyAsync(e.prog_host, e.prog_dev, sizeof(Progress), cudaMemcpyDeviceToHost, e.stream));
      CK(cudaStreamSynchronize(e.stream));
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, e.ev0, e.ev1));
      total_ms += ms;
      j.pending_head += size_t(ne) * j.dev.n;
      j.consumed += uint64_t(ne) * j.dev.n;
      e.prog_host[0].it_outer = 1;
      left -= ne;
    }
    e.seconds_solver += total_ms * 1e-3;
    if (device_ms) *device_ms = total_ms;
    return SGDNET_OK;
  });
}

int sgdnet_session_fit_lambda(sgdnet_session* s, int32_t lambda_ind, sgdnet_rng* rng, uint32_t* epochs, int32_t* converged) {
  if (!s || !rng) return fail(SGDNET_ERR_ARG, "bad session arguments");
  return guarded([&]() -> int {
    Engine& e = s->eng;
    FitJob& j = e.jobs[0];
    if (lambda_ind != e.prog_host[0].lambda_ind) return fail(SGDNET_ERR_ARG, "lambdas must be fitted in path order");
    j.rng = rng;
    e.run(lambda_ind);
    e.settle_rng(j);
    j.pending.clear();
    j.pending_head = 0;
    j.generated = j.consumed;
    j.marks.clear();
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
    // measurement entry: force the deviance + rescale pass for lambda_ind on the current state
    e.prog_host[0].lambda_ind = lambda_ind;
    e.prog_host[0].status = kLambdaDone;
    CK(cudaMemcpyAsync(e.prog_dev, e.prog_host, sizeof(Progress), cudaMemcpyHostToDevice, e.stream));
    CK(cudaEventRecord(e.ev0, e.stream));
    CK(launch_finish_lambda(1, e.fits_dev, e.prog_dev, e.loss_blocks, e.stream));
    e.launches += 2;
    CK(cudaEventRecord(e.ev1, e.stream));
    CK(cudaMemcpyAsync(e.prog_host, e.prog_dev, sizeof(Progress), cudaMemcpyDeviceToHost, e.stream));
    CK(cudaStreamSynchronize(e.stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e.ev0, e.ev1));
    e.seconds_dev += ms * 1e-3;
    if (device_ms) *device_ms = ms;
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
