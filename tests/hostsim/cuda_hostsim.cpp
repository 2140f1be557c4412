// TEST INFRASTRUCTURE ONLY -- see cuda_hostsim.h.
#include "cuda_hostsim.h"
#include <stdio.h>
#include <stdint.h>
#include <mutex>

uint3_ threadIdx, blockIdx;
dim3 blockDim, gridDim;
unsigned char* hostsim_dyn_smem = nullptr;

// Fiber switch.  glibc's swapcontext saves and restores the signal mask with two system calls per switch, and an emulated kernel
// switches fibers at every __syncthreads() / shuffle of every thread: on x86-64 a 14-instruction switch of the callee-saved
// registers and the stack pointer replaces it (the fibers never touch signal masks or the FP control words).  Other
// architectures keep ucontext.
#if defined(__x86_64__)
#define HOSTSIM_FAST_SWITCH 1
extern "C" void hostsim_switch(void** save_sp, void* const* load_sp);
asm(R"(
.text
.globl hostsim_switch
.type hostsim_switch,@function
hostsim_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq (%rsi), %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size hostsim_switch,.-hostsim_switch
)");
#endif

namespace hostsim {
namespace {
#ifdef HOSTSIM_FAST_SWITCH
struct Fiber { void* sp; char* stack; bool done; };
void* sched_sp = nullptr;
#else
struct Fiber { ucontext_t ctx; char* stack; bool done; };
ucontext_t sched_ctx;
#endif
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
#ifdef HOSTSIM_FAST_SWITCH
void trampoline() {
  (*cur_body)();
  fibers[cur].done = true;
  hostsim_switch(&fibers[cur].sp, &sched_sp);
  __builtin_trap();                                  // a finished fiber is never resumed
}
#else
void trampoline() {
  (*cur_body)();
  fibers[cur].done = true;
  swapcontext(&fibers[cur].ctx, &sched_ctx);
}
#endif
}  // namespace

void yield_barrier() {
  int me = cur;
#ifdef HOSTSIM_FAST_SWITCH
  hostsim_switch(&fibers[me].sp, &sched_sp);
#else
  swapcontext(&fibers[me].ctx, &sched_ctx);
#endif
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
          fibers[t].stack = stack_pool[t];
          fibers[t].done = false;
#ifdef HOSTSIM_FAST_SWITCH
          // initial frame: six callee-saved registers, then the "return address" the first switch returns into (the trampoline),
          // then a null return address for the trampoline itself; the trampoline starts with rsp = 8 (mod 16) like any callee
          uintptr_t top = (reinterpret_cast<uintptr_t>(fibers[t].stack) + STACK) & ~(uintptr_t)15;
          void** sp = reinterpret_cast<void**>(top);
          *--sp = nullptr;
          *--sp = reinterpret_cast<void*>(&trampoline);
          for (int r = 0; r < 6; ++r) *--sp = nullptr;
          fibers[t].sp = sp;
#else
          getcontext(&fibers[t].ctx);
          fibers[t].ctx.uc_stack.ss_sp = fibers[t].stack;
          fibers[t].ctx.uc_stack.ss_size = STACK;
          fibers[t].ctx.uc_link = &sched_ctx;
          makecontext(&fibers[t].ctx, (void (*)())trampoline, 0);
#endif
        }
        int live = nthreads;
        while (live > 0) {
          live = 0;
          for (int t = 0; t < nthreads; ++t) {
            if (fibers[t].done) continue;
            cur = t; set_tid(t);
#ifdef HOSTSIM_FAST_SWITCH
            hostsim_switch(&sched_sp, &fibers[t].sp);
#else
            swapcontext(&sched_ctx, &fibers[t].ctx);
#endif
            if (!fibers[t].done) ++live;
          }
        }
      }
  cur = -1;
}
}  // namespace hostsim
