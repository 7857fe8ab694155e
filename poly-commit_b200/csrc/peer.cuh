// Multi-GPU building blocks over NVLink peer memory: one process per GPU, every rank owns one
// "window" (a cudaMalloc'd buffer exported with cudaIpcGetMemHandle) that all other ranks map with cudaIpcOpenMemHandle.
// Kernels of one rank then store straight into the other ranks' HBM (NVLink P2P stores through the NVSwitch) and the ranks
// synchronise with monotonically increasing epoch flags that live in the same windows -- no NCCL call on the data path.
//
//   * MSM sharded by index range (partitioning B): the tail of every rank's Pippenger pipeline PUSHES its bit-plane sums
//     (a 32-byte header + S*c XYZZ points, ~3 KB) into slot [rank] of every peer's window, raises its flag there, waits for
//     the peers' flags in its own window and reads the world_size records back in ONE device-to-host copy: the "point-sum"
//     collective is fused into the pipeline's last kernel (MsmPeerPushBody) instead of an all-gather after it.
//   * four-step NTT sharded by columns / rows: pass 1 stores each element into the row buffer of the rank that owns its row
//     (ntt.cuh NttBlockPeerBody), then PeerSignalBody / PeerWaitBody replace the barrier before pass 2.
//
// Memory-ordering contract: a writer finishes its payload stores, executes __threadfence_system(), then stores the flag
// (volatile, system scope); a reader spins on the flag in ITS OWN memory with volatile loads, then executes
// __threadfence_system() before touching the payload.  Spins are bounded (clock64 budget): a missing peer yields an error
// code, never a hung GPU.
#pragma once
#include "ec.cuh"
#include "msm.cuh"
#include "rt.cuh"
#ifdef PCGPU_EMUL
#include <chrono>
#include <thread>
#endif

namespace pcgpu {

enum { PEER_MAX_WORLD = 16, PEER_RECORD_BYTES = 16384, PEER_FLAG_OFFSET = PEER_MAX_WORLD * PEER_RECORD_BYTES,
       PEER_WINDOW_BYTES = PEER_FLAG_OFFSET + 4096 };
// window layout: record[r] at r * PEER_RECORD_BYTES (written by rank r), flags[r] (uint64 epoch, written by rank r) at
// PEER_FLAG_OFFSET + 8 r, NTT flags at PEER_FLAG_OFFSET + 1024 + 8 r

struct PeerRecordHeader { uint32_t np, S, c, h_split, err, pad[3]; };   // 32 bytes, then np XYZZ points

PCGPU_DEV void peer_store_flag(uint64_t *p, uint64_t v) {
#ifdef __CUDA_ARCH__
  __threadfence_system();
  *reinterpret_cast<volatile uint64_t *>(p) = v;
#else
  *p = v;
#endif
}
PCGPU_DEV uint64_t peer_load_flag(const uint64_t *p) {
#ifdef __CUDA_ARCH__
  return *reinterpret_cast<const volatile uint64_t *>(p);
#else
  return *p;
#endif
}

// Block d copies this rank's record (header + planes, `words` 32-bit words, 16-byte multiple) into slot `rank` of peer d's
// window and raises flag [rank] there with `epoch`.  The header's err word is taken from the pipeline's error word.
struct MsmPeerPushBody {
  const uint32_t *planes; uint32_t plane_words; PeerRecordHeader hdr; const uint32_t *d_err;
  char *win[PEER_MAX_WORLD]; uint32_t rank, world; uint64_t epoch;
  PCGPU_KERNEL_DEV void operator()(size_t blk, uint32_t *) const {
    if (blk >= world) return;
    uint32_t *dst = reinterpret_cast<uint32_t *>(win[blk] + (size_t)rank * PEER_RECORD_BYTES);
    PCGPU_BLOCK_FOR(i, 8) {
      const uint32_t h[8] = {hdr.np, hdr.S, hdr.c, hdr.h_split, d_err ? *d_err : 0u, 0u, 0u, 0u};
      dst[i] = h[i];
    }
    u32x4 *d4 = reinterpret_cast<u32x4 *>(dst + 8);
    const u32x4 *s4 = reinterpret_cast<const u32x4 *>(planes);
    PCGPU_BLOCK_FOR(i, plane_words / 4) { d4[i] = s4[i]; }
#ifdef __CUDA_ARCH__
    __threadfence_system();
#endif
    PCGPU_BLOCK_SYNC();
    PCGPU_BLOCK_FOR(i, 1) { peer_store_flag(reinterpret_cast<uint64_t *>(win[blk] + PEER_FLAG_OFFSET) + rank, epoch); }
  }
};

// raise flag [rank] (at byte offset flag_off inside the windows) on every peer
struct PeerSignalBody {
  char *win[PEER_MAX_WORLD]; uint32_t rank, world, flag_off; uint64_t epoch;
  PCGPU_KERNEL_DEV void operator()(size_t t) const {
    if (t < world) peer_store_flag(reinterpret_cast<uint64_t *>(win[t] + flag_off) + rank, epoch);
  }
};

// wait until every flag of the local window has reached `epoch`; *timed_out |= 1 when the cycle budget runs out first
struct PeerWaitBody {
  const char *local; uint32_t world, flag_off; uint64_t epoch; long long budget_cycles; uint32_t *timed_out;
  PCGPU_KERNEL_DEV void operator()(size_t t) const {
    if (t >= world) return;
    const uint64_t *f = reinterpret_cast<const uint64_t *>(local + flag_off) + t;
#ifdef __CUDA_ARCH__
    const long long t0 = clock64();
    while (peer_load_flag(f) < epoch) {
      if (clock64() - t0 > budget_cycles) { rt::atomic_or(timed_out, 1u); break; }
      __nanosleep(200);
    }
    __threadfence_system();
#elif defined(PCGPU_EMUL)
    // emulation: the "ranks" are host threads of one process; bounded wall-clock spin
    const auto t0 = std::chrono::steady_clock::now();
    long budget_ms = 20000;
    if (const char *e = getenv("PCGPU_EMUL_PEER_WAIT_MS")) { long v = atol(e); if (v > 0) budget_ms = v; }
    while (*reinterpret_cast<const volatile uint64_t *>(f) < epoch) {
      if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(budget_ms)) { rt::atomic_or(timed_out, 1u); break; }
      std::this_thread::yield();
    }
    (void)budget_cycles;
#else
    (void)f;
#endif
  }
};

}  // namespace pcgpu
