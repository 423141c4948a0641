// Runtime of the CUDA execution-model emulator (see cuda_emul.h).  TEST ONLY.
#include "cuda_emul.h"

namespace emul {

thread_local Fiber* cur = nullptr;
thread_local void* sched_sp = nullptr;
static thread_local const std::function<void()>* cur_body = nullptr;

extern "C" void emul_switch(void** save_sp, void* new_sp);
asm(R"(
.text
.globl emul_switch
.type emul_switch,@function
emul_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emul_switch,.-emul_switch
)");

void yield() {
  Fiber* f = cur;
  emul_switch(&f->sp, sched_sp);
}

static void fiber_entry() {
  (*cur_body)();
  Fiber* f = cur;
  f->done = true;
  f->block->alive--;
  // a thread that exits no longer takes part in barriers
  Block* b = f->block;
  if (b->alive > 0 && b->arrived >= b->alive) { b->arrived = 0; b->gen++; }
  Warp* w = &b->warps[f->lin >> 5];
  w->nlanes--;
  if (w->nlanes > 0 && w->arrived >= w->nlanes) { w->arrived = 0; w->gen++; }
  emul_switch(&f->sp, sched_sp);
  abort();
}

static const size_t kStack = 96 * 1024;
struct StackPool {
  std::vector<char*> stacks;
  ~StackPool() { for (char* s : stacks) free(s); }
  char* get(size_t i) {
    while (stacks.size() <= i) {
      void* p = nullptr;
      if (posix_memalign(&p, 64, kStack) != 0) abort();
      stacks.push_back(static_cast<char*>(p));
    }
    return stacks[i];
  }
};
static thread_local StackPool pool;
static thread_local std::vector<unsigned char> dyn_buf;

static void prepare_fiber(Fiber& f, char* stack) {
  f.stack = stack;
  memset(stack, 0xCD, 64);  // canary
  uintptr_t top = (reinterpret_cast<uintptr_t>(stack) + kStack) & ~uintptr_t(15);
  void** s = reinterpret_cast<void**>(top);
  s[-1] = nullptr;
  s[-2] = reinterpret_cast<void*>(&fiber_entry);
  for (int i = 3; i <= 8; ++i) s[-i] = nullptr;
  f.sp = &s[-8];
  f.done = false;
}

static void init_block(Block& b, unsigned lin_block, dim3 grid, dim3 block, size_t stack_base) {
  b.gdim = grid;
  b.bdim = block;
  b.bid.x = lin_block % grid.x;
  b.bid.y = (lin_block / grid.x) % grid.y;
  b.bid.z = lin_block / (grid.x * grid.y);
  b.nthreads = b.alive = block.x * block.y * block.z;
  b.arrived = 0;
  b.gen = 0;
  b.fibers.assign(b.nthreads, Fiber());
  int nw = (b.nthreads + 31) / 32;
  b.warps.assign(nw, Warp());
  for (int w = 0; w < nw; ++w) b.warps[w].nlanes = std::min(32, b.nthreads - 32 * w);
  for (int t = 0; t < b.nthreads; ++t) {
    Fiber& f = b.fibers[t];
    f.block = &b;
    f.lin = t;
    f.tid.x = t % block.x;
    f.tid.y = (t / block.x) % block.y;
    f.tid.z = t / (block.x * block.y);
    prepare_fiber(f, pool.get(stack_base + t));
  }
}

static void run_fibers(std::vector<Block*>& blocks) {
  size_t remaining = 0;
  for (Block* b : blocks) remaining += b->nthreads;
  size_t idle_rounds = 0;
  while (remaining) {
    size_t before = remaining;
    for (Block* b : blocks) {
      if (b->alive == 0) continue;
      for (Fiber& f : b->fibers) {
        if (f.done) continue;
        cur = &f;
        emul_switch(&sched_sp, f.sp);
        if (f.done) {
          --remaining;
          for (int i = 0; i < 64; ++i)
            if (static_cast<unsigned char>(f.stack[i]) != 0xCD) { fprintf(stderr, "emul: fiber stack overflow\n"); abort(); }
        }
      }
    }
    idle_rounds = (remaining == before) ? idle_rounds + 1 : 0;
    if (idle_rounds > 200000000ull) { fprintf(stderr, "emul: deadlock suspected\n"); abort(); }
  }
  cur = nullptr;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body, bool cooperative) {
  const unsigned nblocks = grid.x * grid.y * grid.z;
  const size_t nthreads = size_t(block.x) * block.y * block.z;
  if (nblocks == 0 || nthreads == 0) return;
  if (cooperative) {
    cur_body = &body;
    std::vector<Block> blocks(nblocks);
    std::vector<Block*> ptrs;
    std::vector<std::vector<unsigned char>> smems(nblocks);
    for (unsigned i = 0; i < nblocks; ++i) {
      init_block(blocks[i], i, grid, block, size_t(i) * nthreads);
      smems[i].assign(smem + 256, 0);
      blocks[i].dyn_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smems[i].data()) + 127) & ~uintptr_t(127));
      ptrs.push_back(&blocks[i]);
    }
    run_fibers(ptrs);
    return;
  }
#pragma omp parallel
  {
    cur_body = &body;
    if (dyn_buf.size() < smem + 256) dyn_buf.assign(smem + 256, 0);
    Block b;
#pragma omp for schedule(dynamic, 1)
    for (long i = 0; i < long(nblocks); ++i) {
      init_block(b, unsigned(i), grid, block, 0);
      b.dyn_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn_buf.data()) + 127) & ~uintptr_t(127));
      std::vector<Block*> one{&b};
      run_fibers(one);
    }
  }
}

}  // namespace emul
