// TEST INFRASTRUCTURE ONLY -- see cuda_hostsim.h.
#include "cuda_hostsim.h"
#include <stdio.h>
#include <mutex>

uint3_ threadIdx, blockIdx;
dim3 blockDim, gridDim;
unsigned char* hostsim_dyn_smem = nullptr;

namespace hostsim {
namespace {
struct Fiber { ucontext_t ctx; char* stack; bool done; };
ucontext_t sched_ctx;
std::vector<Fiber> fibers;
int cur = -1;
const std::function<void()>* cur_body = nullptr;
unsigned long long xchg[1024];
const size_t STACK = 256 * 1024;
std::vector<char*> stack_pool;

void set_tid(int t) {
  threadIdx.x = t % blockDim.x;
  threadIdx.y = (t / blockDim.x) % blockDim.y;
  threadIdx.z = t / (blockDim.x * blockDim.y);
}
void trampoline() {
  (*cur_body)();
  fibers[cur].done = true;
  swapcontext(&fibers[cur].ctx, &sched_ctx);
}
}  // namespace

void yield_barrier() {
  int me = cur;
  swapcontext(&fibers[me].ctx, &sched_ctx);
  // resumed: scheduler has restored cur / threadIdx
}

unsigned long long shfl_u64(unsigned long long v, int src_lane) {
  int me = cur;
  xchg[me] = v;
  yield_barrier();
  int base = me & ~31;
  int src = base + src_lane;
  int nthreads = (int)fibers.size();
  unsigned long long r = (src < nthreads) ? xchg[src] : v;
  yield_barrier();
  return r;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  // one emulated kernel at a time: the fiber scheduler's state is global, and the frame pipeline's tracker thread launches
  // kernels while the main thread enqueues the next frame
  static std::mutex launch_mutex;
  std::lock_guard<std::mutex> guard(launch_mutex);
  static std::vector<unsigned char> dyn;
  if (dyn.size() < smem + 16) dyn.resize(smem + 16);
  hostsim_dyn_smem = dyn.data();
  gridDim = grid; blockDim = block;
  int nthreads = block.x * block.y * block.z;
  while ((int)stack_pool.size() < nthreads) stack_pool.push_back((char*)malloc(STACK));
  cur_body = &body;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
        fibers.assign(nthreads, Fiber());
        for (int t = 0; t < nthreads; ++t) {
          getcontext(&fibers[t].ctx);
          fibers[t].stack = stack_pool[t];
          fibers[t].ctx.uc_stack.ss_sp = fibers[t].stack;
          fibers[t].ctx.uc_stack.ss_size = STACK;
          fibers[t].ctx.uc_link = &sched_ctx;
          fibers[t].done = false;
          makecontext(&fibers[t].ctx, (void (*)())trampoline, 0);
        }
        int live = nthreads;
        while (live > 0) {
          live = 0;
          for (int t = 0; t < nthreads; ++t) {
            if (fibers[t].done) continue;
            cur = t; set_tid(t);
            swapcontext(&sched_ctx, &fibers[t].ctx);
            if (!fibers[t].done) ++live;
          }
        }
      }
  cur = -1;
}
}  // namespace hostsim
