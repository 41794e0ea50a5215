// Per-launch device scratch (the "todo" flags a row-column kernel leaves for the per-pixel kernel behind it).
//
// One allocation per device, made at the first launch that needs it: SLOTS slots of equal capacity,
// handed out in ring order.  A slot carries an event recorded behind the last kernel that uses it; a later
// launch that is given the same slot makes its stream wait for that event first, so any number of launches
// may be in flight on any number of streams (no host-side wait, no allocation in the steady state).
// acquire .. release brackets the enqueue of one launch's kernels and holds the module's mutex, so that the
// event of a slot is always recorded before the slot can be handed out again.
// A launch with more subgrids than a slot holds gets a new, larger pool; the old one is kept until the
// process ends (kernels may still be using it, and cudaFree would synchronise the device), which bounds the
// waste by the geometric growth.
#include <mutex>
#include <vector>

#include "kernels.h"

namespace idgb200 {

namespace {

constexpr int SLOTS = 32;
constexpr size_t MIN_SLOT_INTS = 1u << 16;

struct Pool {
  int *base = nullptr;
  size_t slot_ints = 0;
  cudaEvent_t ev[SLOTS] = {};
  bool used[SLOTS] = {};
  unsigned next = 0;
};

std::mutex g_mu;
std::vector<Pool *> g_pools[64];   // per device, the last one is current

}  // namespace

cudaError_t scratch_acquire(size_t ints, cudaStream_t stream, ScratchLease *lease) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  std::unique_lock<std::mutex> lock(g_mu);
  Pool *p = g_pools[dev].empty() ? nullptr : g_pools[dev].back();
  if (!p || p->slot_ints < ints) {
    Pool *q = new Pool;
    q->slot_ints = ints < MIN_SLOT_INTS ? MIN_SLOT_INTS : ints + ints / 2;
    e = cudaMalloc(&q->base, q->slot_ints * SLOTS * sizeof(int));
    for (int i = 0; i < SLOTS && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&q->ev[i], cudaEventDisableTiming);
    if (e != cudaSuccess) {
      if (q->base) cudaFree(q->base);
      delete q;
      cudaGetLastError();
      return e;
    }
    g_pools[dev].push_back(q);
    p = q;
  }
  const int slot = (int)(p->next++ % SLOTS);
  if (p->used[slot]) {
    e = cudaStreamWaitEvent(stream, p->ev[slot], 0);
    if (e != cudaSuccess) return e;
  }
  lease->ptr = p->base + (size_t)slot * p->slot_ints;
  lease->pool = p;
  lease->slot = slot;
  lock.release();   // held until scratch_release
  return cudaSuccess;
}

cudaError_t scratch_release(const ScratchLease &lease, cudaStream_t stream) {
  Pool *p = static_cast<Pool *>(lease.pool);
  p->used[lease.slot] = true;
  const cudaError_t e = cudaEventRecord(p->ev[lease.slot], stream);
  g_mu.unlock();
  return e;
}

}  // namespace idgb200
