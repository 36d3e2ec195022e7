// trace.cuh — the ray-query kernel (sm_100a): persistent warps that drain a ray stream through the
// node / triangle steps of trace_ray.cuh.
//
//   * work distribution: every warp starts on the chunk with its own global index and then claims 32-ray chunks
//     of the rest of the stream from one global cursor (no queue on the atomic when a launch starts); the
//     chunk's eight input arrays (p, wi, d, flags) are staged into shared memory with TMA bulk copies
//     (cp.async.bulk + mbarrier), double-buffered so the next chunk lands while the current one is
//     being traversed; a partial tail chunk (or a misaligned stream) is staged with plain loads;
//   * per-lane refill: a lane whose ray is finished keeps its result until enough lanes are without
//     work, then all of them write their hit records and take the next rays of the warp's chunk
//     (ballot + popc ranks, no atomics), so a long ray never holds 31 finished lanes hostage;
//   * warp-voted phases: two ballots per iteration say what every lane wants next; the warp then runs
//     either one node step (8 quantised child boxes, 240 instructions) for the lanes that need one,
//     or one triangle step (up to two Möller–Trumbore tests of the open leaf piece, both 48-byte
//     fetches in flight together, ~175 instructions) for the lanes with pending leaf triangles — the
//     vote is weighted towards the cheaper triangle step — so both inner loops run converged instead
//     of every lane dragging the warp through its own leaf loop;
//   * hit record = the index of the triangle held so far; its ids (and, on an exact tie in t only, its
//     tie-break order) are read back from the packed triangle when needed: three registers fewer;
//   * traversal stack: kSmemStack entries per ray in shared memory ([entry][thread], conflict-free,
//     stack pointer in a register); a second instantiation (kDeep) with a local-memory spill tier serves
//     trees deeper than that (the re-grouped trees are 8-10 levels deep and never need it).
// Measured alternatives that did not pay off (profiles/r01_summary.md): L1 prefetch of the next node,
// 64 registers / 8 CTAs per SM, 80 registers / 6 CTAs, sorting the ray stream, fused leaves, other vote biases /
// refill thresholds, 64-ray chunks, shorter claims near the end of the stream, packed FFMA2 plane evaluation,
// plane bytes decoded on the XU / FMA pipes, smaller shared-memory stacks, the stack base pinned in a register,
// a per-group distance bound (costed on the host emulation, not built); round 2 (profiles/r02_summary.md): guided
// (shrinking) claims near the end of the stream (-3..5 %: small chunks starve the refill), handing the second triangle of a
// pair to a lane with nothing to test through shared memory (one Moeller-Trumbore sequence per step instead of two:
// -3.5 % coherent, -4.6 % bounce rays — the exchange costs more issue slots than the second test); prefetch.global.L1 of
// the first triangle of the nearest hit leaf / of the next node at the end of a node step (the two fetches are 22 % of the
// stall samples, profiles/r02_trace_config3_bounce_raw.csv): -5..7 %, every extra instruction costs more than the latency
// it hides; 8 CTAs per SM at 64 registers: -2..3 %.  On the eager-commit loop (profiles/r02_sweep_eager_commit.log,
// r02_sweep_eager_knobs.log): prefetch.global.L1 / .L2 of the next node where its index is committed, a whole loop
// iteration before the loads: -1..2.5 % (with a second prefetch 64 bytes on: -35 % on the terrain); L1::evict_last on node
// loads, L1::no_allocate / evict_first on triangle loads: -0..19 %; not staging the next chunk ahead near the end of the
// stream (claim at half of the current chunk / on demand): loses what the eager commit gained; vote bias 2 / 4, refill at
// 4 / 8 / 10 idle lanes, sharing from 8 / 16 idle lanes: within the noise of the defaults; 16- / 24-ray chunks: -1..4 %
// (profiles/r02_sweep_eager_chunks.log); hit acceptance as predicated moves instead of a branch: -0.5..1.5 % (the branch is
// warp-uniformly not taken on most triangle tests; profiles/r02_sweep_branchfree_accept.log); deferred leaves — a lane
// with pending leaf triangles keeps taking part in node steps, the leaves those hit wait as a second set in shared
// memory: the warp-level simulator (tools/warp_sim, which reproduces this kernel's step counters) predicts 5-7 % fewer
// issue slots per ray, the build does the predicted visits and is 0.5-7 % SLOWER (+112 SASS instructions, a spill slot;
// profiles/r02_warp_sim_deferred_leaves.log).
#pragma once
#include "trace_ray.cuh"

namespace phos {

// tuning knobs (defaults chosen by tools/sweep.py runs on the B200, see profiles/)
#ifndef PHOS_CHUNK
#define PHOS_CHUNK 32
#endif
#ifndef PHOS_REFILL_MIN
#define PHOS_REFILL_MIN 6
#endif
#ifndef PHOS_TRI_BIAS
#define PHOS_TRI_BIAS 3
#endif
#ifndef PHOS_MIN_BLOCKS
#define PHOS_MIN_BLOCKS 7
#endif
#ifndef PHOS_TRI_PAIR
#define PHOS_TRI_PAIR 1
#endif
// The end of a launch (profiles/r02_tail_probe.md).  Guided claims: once fewer than PHOS_TAIL_DIV full chunks per warp
// are unclaimed, a claim takes its share of what is left (>= PHOS_TAIL_MIN rays, a multiple of 4 so that chunk starts
// stay 16-byte aligned for the bulk copies) — no warp sits on two staged chunks while others run dry.  0 = fixed claims.
#ifndef PHOS_TAIL_DIV
#define PHOS_TAIL_DIV 0
#endif
#ifndef PHOS_TAIL_MIN
#define PHOS_TAIL_MIN 4
#endif
// Work sharing: once a warp has nothing left to refill from, lanes without a ray take pending sibling groups off the
// traversal stacks of the lanes that still have one (the same ray, a disjoint part of the tree) and hand their hit
// back when done — the longest ray of a warp no longer runs on one lane while 31 wait.  0 = off.
#ifndef PHOS_TAIL_SHARE
#define PHOS_TAIL_SHARE 1
#endif
#ifndef PHOS_SHARE_MIN_IDLE
#define PHOS_SHARE_MIN_IDLE 12  // lanes without a ray before a round of sharing starts (1: -3 % on coherent rays, 12: -1.4 %; profiles/r02_tail.md)
#endif
#ifndef PHOS_SHARE_COHERENCE_TEST
#define PHOS_SHARE_COHERENCE_TEST 1
#endif
#ifndef PHOS_SHARE_EVERY
#define PHOS_SHARE_EVERY 1      // a round of sharing is tried every n-th iteration of the drain loop
#endif
#ifndef PHOS_SHARE_MIN_UNITS
#define PHOS_SHARE_MIN_UNITS 2
#endif
// The end of the stream: once fewer than PHOS_LAZY_TAIL chunks per warp are unclaimed, a warp stages the chunk after the
// current one only when half of the current one is taken — when the cursor runs dry it then holds part of ONE chunk, not
// one and a half on average.  0 = always stage ahead.  (One claim site in the loop either way: a first version with a
// third inlined `claim` grew the kernel from 2288 to 2960 instructions and lost 4 % on every stream.)  Measured on two
// boxes (profiles/r02_sweep_lazy_staging.log): windows of 4-8 chunks give +0.5..2.5 % on the bounce stream, +0.5..1 % on
// the shadow stream, -0.3..0.5 % on coherent rays — inside the box noise; off by default.
#ifndef PHOS_LAZY_TAIL
#define PHOS_LAZY_TAIL 0
#endif
constexpr uint32_t kNoNode = 0xffffffffu;     // "no node left to test" (Ray traversal state)
constexpr int kChunk = PHOS_CHUNK;           // rays per claimed chunk
constexpr int kRefillMin = PHOS_REFILL_MIN;  // idle lanes that trigger a refill
constexpr int kTraceWarps = kTraceBlock / 32;
constexpr uint32_t kHelper = 0x80000000u;    // Ray::flags of a lane that helps another lane's ray (never written out)

struct TraceArgs {
  phos_rays rays;  // device pointers
  unsigned long long n;
  DevAccel accel;
  unsigned long long* cursor;    // chunk counter (zeroed before launch)
  // kCount only, [8]: nodes box-tested, triangles tested (per lane); warp-level node steps, triangle steps; lanes that took
  // part in the node steps, in the triangle steps; loop iterations; rays traced
  unsigned long long* counters;
  int tma_ok;                    // all eight input arrays are 16-byte aligned
  const uint32_t* n_ptr;         // when set, the stream length is read from HBM (wavefront queues)
#ifdef PHOS_TAIL_PROBE
  unsigned long long* probe;     // tuning probe (tools/tail_probe.py): 5 words per warp, see the end of trace_kernel
#endif
};

#ifdef PHOS_TAIL_PROBE
__device__ __forceinline__ unsigned long long probe_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// kDeep: the tree may be deeper than the shared-memory stack (upload knows the depth).  Only that instantiation
// carries the local-memory spill tier; the common one has no stack-range checks at all in push / pop.
template <bool kCount, bool kDeep>
__global__ void __launch_bounds__(kTraceBlock, PHOS_MIN_BLOCKS) trace_kernel(const TraceArgs P) {
  __shared__ uint2 s_stack[kSmemStack * kTraceBlock];
  __shared__ alignas(128) uint32_t s_stage[kTraceWarps][2][8][kChunk];
  __shared__ alignas(8) unsigned long long s_bar[kTraceWarps][2];
  __shared__ int s_share_min[kTraceWarps];

  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#ifdef PHOS_TAIL_PROBE
  const unsigned long long pr_start = probe_now();
  unsigned long long pr_dry = 0, pr_iters = 0, pr_lanes = 0;
#endif
  const unsigned long long N = P.n_ptr ? (unsigned long long)*P.n_ptr : P.n;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint32_t(*stage)[8][kChunk] = s_stage[warp];
  unsigned long long* bar = s_bar[warp];
  if (lane == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();

  const uint32_t* src[8] = {(const uint32_t*)P.rays.px, (const uint32_t*)P.rays.py, (const uint32_t*)P.rays.pz,
                            (const uint32_t*)P.rays.wx, (const uint32_t*)P.rays.wy, (const uint32_t*)P.rays.wz,
                            (const uint32_t*)P.rays.d,  (const uint32_t*)P.rays.flags};

  // ---- the warp's view of the stream (all warp-uniform) ---------------------------------------------
  uint32_t phase = 0;  // bit b: parity the next wait on buffer b expects
  int cur_buf = 0;
  uint32_t cur_base = 0, nxt_base = 0;  // ray indices fit 32 bits: launch_trace splits longer streams
  uint32_t cur_cnt = 0, taken = 0, nxt_cnt = 0;
  bool nxt_tma = false, exhausted = false;
  bool lazy = false;  // PHOS_LAZY_TAIL: the current chunk lies in the last part of the stream

  bool first_claim = true;
  auto claim = [&](int buf) {  // claim the next chunk of the stream and start staging it into `buf`
    nxt_cnt = 0;
    if (exhausted) return;
    unsigned long long base = 0;
    uint32_t want = kChunk;
    if (first_claim) {  // chunk = global warp index: 4144 warps do not queue up on one atomic at the start
      base = ((unsigned long long)blockIdx.x * kTraceWarps + warp) * kChunk;
      first_claim = false;
    } else {
      // (recomputed per claim rather than held in registers over the traversal loop)
      const unsigned long long static_rays = (unsigned long long)gridDim.x * (kTraceWarps * kChunk);  // every warp's first chunk is assigned
      if (lane == 0) {
        if (PHOS_TAIL_DIV) {
          const float tail_share = __fdividef(1.0f, (float)((PHOS_TAIL_DIV ? PHOS_TAIL_DIV : 1) * kTraceWarps * gridDim.x));  // what is unclaimed right now (a plain read just before the atomic), shared out over the warps
          const unsigned long long pos = *(volatile unsigned long long*)P.cursor + static_rays;
          const float left = pos < N ? (float)(uint32_t)(N - pos) : 0.0f;
          want = min((uint32_t)kChunk, max((uint32_t)PHOS_TAIL_MIN, (uint32_t)(left * tail_share) & ~3u));
        }
        base = atomicAdd(P.cursor, (unsigned long long)want);  // the cursor counts rays past the static chunks
      }
      base = __shfl_sync(0xffffffffu, base, 0) + static_rays;
      if (PHOS_TAIL_DIV) want = __shfl_sync(0xffffffffu, want, 0);
    }
    if (base >= N) {
      exhausted = true;
#ifdef PHOS_TAIL_PROBE
      pr_dry = probe_now();
#endif
      return;
    }
    nxt_base = (uint32_t)base;
    nxt_cnt = (uint32_t)min((unsigned long long)want, N - base);
    nxt_tma = P.tma_ok && nxt_cnt == want;  // whole claims are multiples of 4 rays: 16-byte sizes and addresses
    if (nxt_tma) {
      if (lane == 0) {
        mbar_expect_tx(&bar[buf], 8u * nxt_cnt * 4u);
#pragma unroll
        for (int a = 0; a < 8; ++a) bulk_g2s(stage[buf][a], src[a] + base, nxt_cnt * 4u, &bar[buf]);
      }
    } else {
#pragma unroll
      for (int a = 0; a < 8; ++a)
        for (uint32_t k = lane; k < nxt_cnt; k += 32) stage[buf][a][k] = src[a][base + k];
    }
    __syncwarp();
  };
  auto advance = [&]() -> bool {  // the staged chunk becomes current; prefetch the one after
    if (nxt_cnt == 0) return false;
    const int b = cur_buf ^ 1;
    if (nxt_tma) {
      mbar_wait(&bar[b], (phase >> b) & 1u);
      phase ^= 1u << b;
    }
    cur_buf = b;
    cur_base = nxt_base;
    cur_cnt = nxt_cnt;
    taken = 0;
    __syncwarp();  // every lane is done reading the buffer we are about to refill
    if (PHOS_LAZY_TAIL) {
      const unsigned long long window = (unsigned long long)gridDim.x * (kTraceWarps * kChunk * PHOS_LAZY_TAIL);
      lazy = (unsigned long long)cur_base + window >= N;
      nxt_cnt = 0;  // claimed from the refill loop below
    } else {
      claim(b ^ 1);
    }
    return true;
  };
  claim(1);

  // ---- per-lane traversal state ----------------------------------------------------------------------
  // traversal stack: sp in a register, entries in this thread's shared-memory column (plain STS / LDS),
  // deeper entries in a local array that ordinary trees never touch (the Stack struct of trace_ray.cuh
  // is the same thing for the one-ray loop; taking it apart here keeps sp out of local memory)
  uint2* const my_stack = s_stack + threadIdx.x;
  uint2 spill[kDeep ? kSpillStack : 1];
  int sp = 0;
  auto push = [&](uint2 v) {
    if (!kDeep || sp < kSmemStack) my_stack[sp * kTraceBlock] = v;
    else if (sp < kSmemStack + kSpillStack) spill[sp - kSmemStack] = v;
    ++sp;
  };
  auto pop = [&]() -> uint2 {
    --sp;
    return (!kDeep || sp < kSmemStack) ? my_stack[sp * kTraceBlock] : spill[sp - kSmemStack];
  };
  Ray r;
  RayDir rd;
  r.flags = 0;
  rd.oct = 0;
  uint32_t ridx = 0;
  bool has_ray = false;
  // The node this lane tests next.  A node step ENDS by committing the one after it: the nearest pending child of the
  // node just tested (its siblings go onto the stack as one group), or, when all its children missed, the nearest child
  // of the group on top of the stack.  So every pending group lives on the stack, a lane carries one node index instead
  // of a group of siblings, "has node work" is one compare, and a lane without a ray has lt == 0 and node == kNoNode —
  // the two work predicates of the loop head need no has_ray (+3.5..5 % over popping at the start of the next node step,
  // profiles/r02_sweep_eager_commit.log).
  uint32_t node = kNoNode;
  // leaf state: `lt` = hit leaf slots of the current node still to open (key space, bits 0-7) | triangles left
  // in the open leaf (bits 8-11); tptr = next triangle; lcounts / lbase = the node's leaf table
  uint32_t lt = 0, tptr = 0, lcounts = 0, lbase = 0;
  uint32_t n_nodes = 0, n_tris = 0;
  uint32_t w_node = 0, w_tri = 0, l_node = 0, l_tri = 0, w_iter = 0, n_rays = 0;  // kCount: warp-level step statistics

  // hit record of a finished ray -> the stream
  auto retire = [&]() {
    if (r.tri != kNoTri) {  // something was accepted (accept_hit marks shadow rays too)
      P.rays.d[ridx] = r.d;
      P.rays.flags[ridx] = r.flags;
      if (!(r.flags & PHOS_SHADOW)) {
        const uint4 ids = __ldg(P.accel.tris + 3ull * r.tri + 2);  // changed and not SHADOW: a triangle is held
        P.rays.mesh[ridx] = ids.y;
        P.rays.face[ridx] = ids.z;
        P.rays.u[ridx] = r.u;
        P.rays.v[ridx] = r.v;
      }
    }
    has_ray = false;
  };
  // 3. vote: one node step or one triangle step for the whole warp (a triangle step is cheaper than a node step:
  // PHOS_TRI_BIAS weights the vote)
  auto step = [&](const bool tri_work, const bool node_work, const int n_tri, const int n_node) {
    if (kCount) {
      if (PHOS_TRI_BIAS * n_tri >= n_node) {
        ++w_tri;
        l_tri += n_tri;
      } else {
        ++w_node;
        l_node += n_node;
      }
    }
    if (PHOS_TRI_BIAS * n_tri >= n_node) {
      if (tri_work && (lt >> 8) == 0u) {  // open the next hit leaf, nearest octant first
        const uint32_t slot = (__ffs(lt) - 1) ^ rd.oct;
        lt &= lt - 1u;
        lt |= ((lcounts >> (4u * slot)) & 15u) << 8;  // 0 for an empty slot
        tptr = lbase + nibble_prefix(lcounts, slot);
      }
      if (tri_work) {
        if (lt >> 8) {
          const uint4* tp = P.accel.tris + 3ull * tptr;
          const uint4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
#if PHOS_TRI_PAIR
          // two triangles of the leaf per step: both 48-byte fetches are in flight together.  The second
          // triangle's registers are left undefined (not copied from the first) when there is none.
          const bool two = lt >= 0x200u;
          uint4 a2, b2, c2;
          asm("" : "=r"(a2.x), "=r"(a2.y), "=r"(a2.z), "=r"(a2.w));
          asm("" : "=r"(b2.x), "=r"(b2.y), "=r"(b2.z), "=r"(b2.w));
          asm("" : "=r"(c2.x), "=r"(c2.y), "=r"(c2.z), "=r"(c2.w));
          if (two) {
            a2 = __ldg(tp + 3);
            b2 = __ldg(tp + 4);
            c2 = __ldg(tp + 5);
          }
#else
          const bool two = false;
          const uint4 a2 = a, b2 = b, c2 = c;
#endif
          float ds, us, vs;
          bool done = false;
          if (mt_triangle(a, b, c, r.ox, r.oy, r.oz, r.wx, r.wy, r.wz, ds, us, vs) && accept_hit(P.accel, r, ds, us, vs, tptr))
            done = (r.flags & PHOS_SHADOW) != 0u;
          if (two && !done && mt_triangle(a2, b2, c2, r.ox, r.oy, r.oz, r.wx, r.wy, r.wz, ds, us, vs) &&
              accept_hit(P.accel, r, ds, us, vs, tptr + 1u))
            done = (r.flags & PHOS_SHADOW) != 0u;
          // advance past what was tested (the pair flag is re-read from lt: nothing extra stays live over the tests)
          const bool adv2 = PHOS_TRI_PAIR && lt >= 0x200u;
          tptr += adv2 ? 2u : 1u;
          lt -= adv2 ? 0x200u : 0x100u;
          if (kCount) n_tris += adv2 ? 2u : 1u;
          if (done) {  // any-hit: this ray is finished
            lt = 0u;
            node = kNoNode;
            sp = 0;
          }
        }
      }
    } else if (node_work) {
      const NodeHits h = node_test(P.accel, node, r.ox, r.oy, r.oz, rd, r.d * PHOS_CULL_SLACK);
      if (kCount) ++n_nodes;
      lt = h.leaf;
      lcounts = h.counts;
      lbase = h.tri_base;
      uint2 g = make_uint2(h.child_base, h.imask | (h.inner << 8));
      if ((g.y >> 8) == 0u && sp != 0) g = pop();  // its own children all missed: the nearest group still pending
      if (g.y >> 8) {
        node = take_child(g, rd.oct);
        if (g.y >> 8) push(g);
      } else {
        node = kNoNode;
      }
    }
  };

  // ---- phase 1: the stream still has rays for this warp --------------------------------------------------------
  for (;;) {
    if (kCount) ++w_iter;
    // 1. what every lane wants to do next (two ballots drive everything else)
    const bool tri_work = lt != 0u;
    const bool node_work = !tri_work && node != kNoNode;
    const unsigned tl = __ballot_sync(0xffffffffu, tri_work);
    const unsigned nl = __ballot_sync(0xffffffffu, node_work);
    const int n_tri = __popc(tl), n_node = __popc(nl);  // disjoint sets: the rest of the warp is without work
    // 2. enough lanes without work (finished rays or empty lanes): retire and refill from the chunk
    if (32 - n_tri - n_node >= kRefillMin) {
      if (has_ray && !tri_work && !node_work) retire();
      if (taken < cur_cnt || nxt_cnt != 0) {
        unsigned idle = __ballot_sync(0xffffffffu, !has_ray);
        while (idle) {
          if (taken == cur_cnt && !advance()) break;
          const uint32_t rank = __popc(idle & lt_mask);
          const uint32_t take = min((uint32_t)__popc(idle), cur_cnt - taken);
          if (!has_ray && rank < take) {
            const uint32_t k = taken + rank;
            const uint32_t fl = stage[cur_buf][7][k];
            if (!(fl & PHOS_MASKED)) {  // MASKED rays are consumed without being traced
              r.ox = __uint_as_float(stage[cur_buf][0][k]);
              r.oy = __uint_as_float(stage[cur_buf][1][k]);
              r.oz = __uint_as_float(stage[cur_buf][2][k]);
              r.wx = __uint_as_float(stage[cur_buf][3][k]);
              r.wy = __uint_as_float(stage[cur_buf][4][k]);
              r.wz = __uint_as_float(stage[cur_buf][5][k]);
              r.d = __uint_as_float(stage[cur_buf][6][k]);
              r.flags = fl;
              r.tri = kNoTri;
              r.u = r.v = 0.0f;
              rd = make_raydir(r.wx, r.wy, r.wz);
              ridx = cur_base + k;
              node = 0u;  // the root
              sp = 0;
              lt = 0u;
              has_ray = true;
              if (kCount) ++n_rays;
            }
          }
          taken += take;
          if (PHOS_LAZY_TAIL && nxt_cnt == 0 && !exhausted && (!lazy || 2u * taken >= cur_cnt)) claim(cur_buf ^ 1);
          idle = __ballot_sync(0xffffffffu, !has_ray);
        }
        continue;  // new rays: vote again
      }
      if (PHOS_TAIL_SHARE || (tl | nl) == 0u) break;  // nothing left to refill from: phase 2 (or: every lane retired)
    }
    step(tri_work, node_work, n_tri, n_node);
  }

#if PHOS_TAIL_SHARE
  uint32_t share_tick = 0;
  // Coherent streams (camera rays) end warp-wide together: there is little to share and the exchange only costs (-2.5 %
  // on config 2, profiles/r02_summary.md).  A warp whose remaining rays all point into ONE octant takes that as the sign
  // of a coherent stream and does not share; bounce and shadow streams mix octants in every warp.
  // (The threshold lives in shared memory: held in a register through phase 2 it cost a spill slot in the HOT loop; a
  // stateless test inside the sharing round and a bool flag compiled to a 4 % slower hot loop,
  // profiles/r02_sweep_share_coherence.log.)
#if PHOS_SHARE_COHERENCE_TEST
  {
    const unsigned with_ray = __ballot_sync(0xffffffffu, has_ray);
    int same = 0;
    if (has_ray) __match_all_sync(with_ray, rd.oct, &same);
    const bool coherent = __any_sync(0xffffffffu, has_ray && same);
    if (lane == 0) s_share_min[warp] = coherent ? 33 : PHOS_SHARE_MIN_IDLE;  // kept in shared memory: no register lives through phase 2 for it
    __syncwarp();
  }
#else
  if (lane == 0) s_share_min[warp] = PHOS_SHARE_MIN_IDLE;
  __syncwarp();
#endif
#define share_min_idle (s_share_min[warp])
  // ---- phase 2: nothing left to refill from — lanes without work help the lanes that still have a ray ---------------
  // A helper holds a copy of the ray (flag bit 31, `ridx` = the lane that owns the ray) and one group of pending siblings
  // taken off the owner's stack (the same ray, a disjoint part of the tree); when it runs out of work its hit is merged
  // into the owner's with the ordinary acceptance rule (closest, ties to the lower order; any-hit: the verdict), and the
  // owner writes the ray once no helper is left.  The longest ray of a warp no longer runs on one lane while 31 wait.
  for (;;) {
    if (kCount) ++w_iter;
    bool tri_work = lt != 0u;
    bool node_work = !tri_work && node != kNoNode;
    unsigned tl = __ballot_sync(0xffffffffu, tri_work);
    unsigned nl = __ballot_sync(0xffffffffu, node_work);
#ifdef PHOS_TAIL_PROBE
    ++pr_iters;
    pr_lanes += __popc(tl | nl);
#endif
    // the block below runs when a lane has finished (hand over / retire) or when enough lanes are free for a round of sharing
    // (the rest of the drain pays two ballots, not the whole protocol)
    const unsigned fin_any = __ballot_sync(0xffffffffu, has_ray && !tri_work && !node_work);
    const unsigned free_now = __ballot_sync(0xffffffffu, !has_ray);
    if (free_now == 0xffffffffu) break;  // every lane retired
    ++share_tick;
    if (fin_any != 0u || (__popc(free_now) >= share_min_idle && (share_tick % PHOS_SHARE_EVERY) == 0u)) {
      // helpers that are done hand their record to the owner
      unsigned fh = __ballot_sync(0xffffffffu, has_ray && !tri_work && !node_work && (r.flags & kHelper));
      bool changed = fh != 0u;
      while (fh) {
        const int h = __ffs(fh) - 1;
        fh &= fh - 1u;
        const uint32_t own = __shfl_sync(0xffffffffu, ridx, h);
        const float hd = __shfl_sync(0xffffffffu, r.d, h), hu = __shfl_sync(0xffffffffu, r.u, h), hv = __shfl_sync(0xffffffffu, r.v, h);
        const uint32_t ht = __shfl_sync(0xffffffffu, r.tri, h);
        if (lane == own && ht != kNoTri && ht != r.tri) {
          if (r.flags & PHOS_SHADOW) {  // any-hit: the helper accepted a triangle — occluded; the owner's own traversal stops
            r.flags |= PHOS_HIT;
            r.d = fminf(r.d, hd);
            r.tri = ht;
            lt = 0u;
            node = kNoNode;
            sp = 0;
          } else {
            accept_hit(P.accel, r, hd, hu, hv, ht);
          }
        }
        if ((int)lane == h) has_ray = false;
      }
      // owners that are done and have no helper left write their ray
      const bool busy = lt != 0u || node != kNoNode;
      const uint32_t group = !has_ray ? 32u + lane : (r.flags & kHelper) ? ridx : lane;
      const unsigned peers = __match_any_sync(0xffffffffu, group);
      const bool leaves = has_ray && !busy && !(r.flags & kHelper) && peers == (1u << lane);
      if (leaves) retire();
      changed = changed || __any_sync(0xffffffffu, leaves);  // (warp-uniform: it guards warp collectives below)
      // free lanes take a group of pending siblings each from the lanes that can spare one: the top entry of the stack,
      // or, with an empty stack, all but the nearest pending child of the current group
      // every pending group is on the stack, and a lane with a stacked group also holds the node it tests next: it gives
      // its top entry and keeps that node (profiles/r02_sweep_eager_knobs.log: idle-lane thresholds 8 / 12 / 16 are equal)
      const bool can_give = sp > 0 && sp + 1 >= PHOS_SHARE_MIN_UNITS;
      const unsigned idle = __ballot_sync(0xffffffffu, !has_ray), don = __ballot_sync(0xffffffffu, can_give);
      if (idle == 0xffffffffu) break;  // every lane retired
      if (__popc(idle) >= share_min_idle && don != 0u) {
        const int pairs = min(__popc(idle), __popc(don));
        const bool take = !has_ray && __popc(idle & lt_mask) < pairs;
        const bool give = can_give && __popc(don & lt_mask) < pairs;
        uint2 g = make_uint2(0u, 0u);
        if (give) g = pop();
        const int from = take ? (int)__fns(don, 0, __popc(idle & lt_mask) + 1) : (int)lane;
        const uint32_t gx = __shfl_sync(0xffffffffu, g.x, from), gy = __shfl_sync(0xffffffffu, g.y, from);
        const float ox = __shfl_sync(0xffffffffu, r.ox, from), oy = __shfl_sync(0xffffffffu, r.oy, from), oz = __shfl_sync(0xffffffffu, r.oz, from);
        const float wx = __shfl_sync(0xffffffffu, r.wx, from), wy = __shfl_sync(0xffffffffu, r.wy, from), wz = __shfl_sync(0xffffffffu, r.wz, from);
        const float dd = __shfl_sync(0xffffffffu, r.d, from), uu = __shfl_sync(0xffffffffu, r.u, from), vv = __shfl_sync(0xffffffffu, r.v, from);
        const uint32_t tt = __shfl_sync(0xffffffffu, r.tri, from), ff = __shfl_sync(0xffffffffu, r.flags, from);
        const uint32_t oo = __shfl_sync(0xffffffffu, (r.flags & kHelper) ? ridx : lane, from);
        if (take) {
          r.ox = ox; r.oy = oy; r.oz = oz;
          r.wx = wx; r.wy = wy; r.wz = wz;
          r.d = dd; r.u = uu; r.v = vv;
          r.tri = tt;
          r.flags = ff | kHelper;
          rd = make_raydir(wx, wy, wz);
          ridx = oo;
          uint2 grp = make_uint2(gx, gy);
          sp = 0;
          node = take_child(grp, rd.oct);
          if (grp.y >> 8) push(grp);
          lt = 0u;
          has_ray = true;
        }
        changed = true;
      }
      if (changed) {  // vote again
        tri_work = lt != 0u;
        node_work = !tri_work && node != kNoNode;
        tl = __ballot_sync(0xffffffffu, tri_work);
        nl = __ballot_sync(0xffffffffu, node_work);
      }
    }
    step(tri_work, node_work, __popc(tl), __popc(nl));
  }
#undef share_min_idle
#endif
#ifdef PHOS_TAIL_PROBE
  if (lane == 0 && P.probe) {
    unsigned long long* o = P.probe + 5ull * (blockIdx.x * kTraceWarps + warp);
    o[0] = pr_start;
    o[1] = pr_dry;
    o[2] = probe_now();
    o[3] = pr_iters;
    o[4] = pr_lanes;
  }
#endif
  if (kCount) {
    atomicAdd(P.counters, (unsigned long long)n_nodes);
    atomicAdd(P.counters + 1, (unsigned long long)n_tris);
    atomicAdd(P.counters + 7, (unsigned long long)n_rays);
    if (lane == 0) {
      atomicAdd(P.counters + 2, (unsigned long long)w_node);
      atomicAdd(P.counters + 3, (unsigned long long)w_tri);
      atomicAdd(P.counters + 4, (unsigned long long)l_node);
      atomicAdd(P.counters + 5, (unsigned long long)l_tri);
      atomicAdd(P.counters + 6, (unsigned long long)w_iter);
    }
  }
}

}  // namespace phos
