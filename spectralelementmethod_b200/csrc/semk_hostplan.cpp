// semk_hostplan.cpp -- host-side patch plan builder (no CUDA calls).
//
// The reference assembles by looping over elements and scatter-adding through
// the L2G map (sem/discrete.py:478-500).  The engine instead groups element
// slots into patches (one CTA each).  Per patch we tabulate the distinct
// global nodes it touches, split into
//   private nodes -- every element containing the node lies in this patch:
//                    the CTA owns the final value and stores it directly;
//   shared nodes  -- touched by more than one patch: the CTA writes its
//                    partial sum to an interface slot and a second, tiny
//                    kernel sums the slots of each shared node in a fixed
//                    order (deterministic, no floating-point atomics).
// Within a patch the CTA assembles by gathering: an inverse table lists, for
// every patch node, where its element-local contributions sit in the CTA's
// scratch, so each node is summed in a fixed order without atomics.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#include <parallel/algorithm>
#endif

#include "../../include/semk.h"

extern void semk_set_error(const std::string &msg);  // semk_api.cu

struct semk_hostplan {
  int n1 = 0;
  int pe = 0;
  int64_t n_elem = 0, n_nodes = 0;
  int64_t scalars[SEMK_PS_COUNT] = {0};
  std::vector<int32_t> patch_node_ptr, patch_npriv, patch_nnodes, patch_slot_base, shared_ptr,
      shared_slot;
  std::vector<uint32_t> pnode, shared_node, pnblk, shared_rec, shared_ext, shared_chunk, patch_hdr,
      patch_maxnode;
  std::vector<int32_t> chunk_maxpatch, rec_maxpatch;
  std::vector<uint16_t> elblk, invblk;
  std::vector<uint16_t> eloc;
  std::vector<int64_t> elem_of_slot;
};

extern "C" int semk_scratch_row_stride(int n1, int elems_per_patch) {
  return ((n1 * elems_per_patch - 1 + 15) & ~15) + 1;  // == scratch_row_stride (semk_apply.cu)
}

namespace {
template <class It, class Cmp>
void semk_plan_sort(It first, It last, int threads, Cmp cmp) {
#if defined(_OPENMP) && defined(_PARALLEL_ALGORITHM)
  if (threads > 1 && last - first > (1 << 16)) {
    const int before = omp_get_max_threads();
    omp_set_num_threads(threads);
    __gnu_parallel::sort(first, last, cmp);
    omp_set_num_threads(before);
    return;
  }
#endif
  (void)threads;
  std::sort(first, last, cmp);
}

// threads of the per-patch passes: the caller's request, else what OpenMP would use
int plan_threads(int requested) {
#ifdef _OPENMP
  int t = requested > 0 ? requested : omp_get_max_threads();
  return t < 1 ? 1 : (t > 256 ? 256 : t);
#else
  (void)requested;
  return 1;
#endif
}
}  // namespace

extern "C" int semk_hostplan_create(int n1, int64_t n_elem, int64_t n_nodes, const uint32_t *l2g,
                                    const int64_t *elem_order, int64_t n_order,
                                    int elems_per_patch, const uint8_t *dirichlet,
                                    semk_hostplan **out) {
  return semk_hostplan_create_mt(n1, n_elem, n_nodes, l2g, elem_order, n_order, elems_per_patch,
                                 dirichlet, 0, out);
}

// The per-patch passes (node lists, index tables, device blocks, inverse tables) run on
// `n_threads` threads; every table is byte-identical to the single-threaded build (patches
// are independent once the private / shared split is known; offsets come from prefix sums).
extern "C" int semk_hostplan_create_mt(int n1, int64_t n_elem, int64_t n_nodes,
                                       const uint32_t *l2g, const int64_t *elem_order,
                                       int64_t n_order, int elems_per_patch,
                                       const uint8_t *dirichlet, int n_threads,
                                       semk_hostplan **out) {
  if (!out) return SEMK_ERR_INVALID;
  const int T = plan_threads(n_threads);
  (void)T;
  *out = nullptr;
  if (n1 < 2 || n1 > SEMK_MAX_N1) {
    semk_set_error("semk_hostplan_create: n1 must be in [2, 17]");
    return SEMK_ERR_UNSUPPORTED;
  }
  if (n_elem < 1 || n_nodes < 1 || !l2g || elems_per_patch < 1) {
    semk_set_error("semk_hostplan_create: bad sizes or null l2g");
    return SEMK_ERR_INVALID;
  }
  if (n_nodes > (int64_t)SEMK_NODE_ID_MASK) {
    semk_set_error("semk_hostplan_create: more than 2^30-1 nodes per rank");
    return SEMK_ERR_UNSUPPORTED;
  }
  const int NN = n1 * n1;
  // eloc block of one patch: [m][le][t], NN*PE uint16 rounded up to 16-byte multiples (TMA)
  const int64_t ES = ((int64_t)NN * elems_per_patch + 7) & ~(int64_t)7;
  const int PE = elems_per_patch;
  if ((int64_t)PE * NN > 65535) {
    semk_set_error("semk_hostplan_create: patch too large for 16-bit local indices");
    return SEMK_ERR_INVALID;
  }
  semk_hostplan *P = new (std::nothrow) semk_hostplan();
  if (!P) return SEMK_ERR_INVALID;
  try {
    P->n1 = n1;
    P->pe = PE;
    P->n_elem = n_elem;
    P->n_nodes = n_nodes;
    // The engine order may contain EMPTY slots (-1): padding that keeps every patch a
    // compact tile when the mesh is not a whole number of tiles.
    if (!elem_order || n_order <= 0) n_order = n_elem;
    if (n_order < n_elem) {
      delete P;
      semk_set_error("semk_hostplan_create: elem_order shorter than the number of elements");
      return SEMK_ERR_INVALID;
    }
    const int64_t n_patch = (n_order + PE - 1) / PE;
    const int64_t n_slot_elems = n_patch * PE;

    // slot -> element
    P->elem_of_slot.resize(n_order);
    {
      std::vector<uint8_t> seen(n_elem, 0);
      int64_t n_seen = 0;
      for (int64_t s = 0; s < n_order; ++s) {
        int64_t e = elem_order ? elem_order[s] : s;
        if (e == -1) {
          P->elem_of_slot[s] = -1;
          continue;
        }
        if (e < 0 || e >= n_elem || seen[e]) {
          delete P;
          semk_set_error("semk_hostplan_create: elem_order is not a permutation");
          return SEMK_ERR_INVALID;
        }
        seen[e] = 1;
        ++n_seen;
        P->elem_of_slot[s] = e;
      }
      if (n_seen != n_elem) {
        delete P;
        semk_set_error("semk_hostplan_create: elem_order is not a permutation");
        return SEMK_ERR_INVALID;
      }
      for (int64_t p = 0; p < n_patch; ++p) {
        bool any = false;
        for (int64_t s = p * PE; s < std::min<int64_t>((p + 1) * PE, n_order); ++s)
          any = any || P->elem_of_slot[s] >= 0;
        if (!any) {
          delete P;
          semk_set_error("semk_hostplan_create: a patch consists of empty slots only");
          return SEMK_ERR_INVALID;
        }
      }
    }
    for (int64_t i = 0; i < n_elem * NN; ++i) {
      if (l2g[i] >= (uint64_t)n_nodes) {
        delete P;
        semk_set_error("semk_hostplan_create: l2g entry out of range");
        return SEMK_ERR_INVALID;
      }
    }

    // pass 1: which nodes are touched by more than one patch (serial: the sweep is bound by
    // first-touching its two n_nodes-sized arrays; a compare-and-swap version on 8 threads
    // was no faster)
    std::vector<int32_t> first_patch(n_nodes, -1);
    std::vector<uint8_t> multi(n_nodes, 0);  // 0 private, 1 shared (interface slots)
    for (int64_t p = 0; p < n_patch; ++p) {
      const int64_t s0 = p * PE, s1 = std::min<int64_t>(s0 + PE, n_order);
      for (int64_t s = s0; s < s1; ++s) {
        if (P->elem_of_slot[s] < 0) continue;
        const uint32_t *row = l2g + P->elem_of_slot[s] * NN;
        for (int k = 0; k < NN; ++k) {
          const uint32_t g = row[k];
          if (first_patch[g] < 0)
            first_patch[g] = (int32_t)p;
          else if (first_patch[g] != (int32_t)p)
            multi[g] = 1;
        }
      }
    }
    std::vector<int32_t>().swap(first_patch);

    // shared node list (ascending id) and slot counts: count per block of nodes, prefix,
    // fill
    std::vector<int32_t> shared_index(n_nodes, -1);
    {
      const int64_t kBlk = 1 << 16;
      const int64_t n_blk = (n_nodes + kBlk - 1) / kBlk;
      std::vector<int64_t> first_of(n_blk + 1, 0);
#pragma omp parallel for num_threads(T) schedule(static)
      for (int64_t b = 0; b < n_blk; ++b) {
        int64_t c = 0;
        const int64_t g1 = std::min<int64_t>((b + 1) * kBlk, n_nodes);
        for (int64_t g = b * kBlk; g < g1; ++g) c += multi[g];
        first_of[b + 1] = c;
      }
      for (int64_t b = 0; b < n_blk; ++b) first_of[b + 1] += first_of[b];
      if (first_of[n_blk] > INT32_MAX) {
        delete P;
        semk_set_error("semk_hostplan_create: shared node count overflows int32");
        return SEMK_ERR_UNSUPPORTED;
      }
      P->shared_node.resize((size_t)first_of[n_blk]);
#pragma omp parallel for num_threads(T) schedule(static)
      for (int64_t b = 0; b < n_blk; ++b) {
        int64_t i = first_of[b];
        const int64_t g1 = std::min<int64_t>((b + 1) * kBlk, n_nodes);
        for (int64_t g = b * kBlk; g < g1; ++g)
          if (multi[g] == 1) {
            shared_index[g] = (int32_t)i;
            uint32_t v = (uint32_t)g | SEMK_NODE_SHARED;
            if (dirichlet && dirichlet[g]) v |= SEMK_NODE_DIRICHLET;
            P->shared_node[(size_t)i++] = v;
          }
      }
    }
    const int64_t n_shared = (int64_t)P->shared_node.size();
    P->shared_ptr.assign(n_shared + 1, 0);

    // pass 2: per-patch tables, one sweep.  Every thread takes a CONTIGUOUS range of patches
    // and appends their node lists / interface-slot owners to its own buffers; the index
    // tables have a fixed stride and are written in place.  Offsets then come from prefix
    // sums over the patches and every thread's buffer is copied to its final position (it is
    // exactly the final sub-array of its patch range).
    P->patch_node_ptr.assign(n_patch + 1, 0);
    P->patch_npriv.assign(n_patch, 0);
    P->patch_nnodes.assign(n_patch, 0);
    P->patch_slot_base.assign(n_patch, 0);
    P->eloc.assign((size_t)n_patch * ES, 0);
    // global node id -> patch-local index, for one patch at a time (open addressing; the
    // single-threaded build used an n_nodes-sized scratch array instead)
    struct NodeMap {
      std::vector<uint32_t> key;
      std::vector<int32_t> val;
      std::vector<uint32_t> used;
      uint32_t mask = 0;
      int shift = 0;
      explicit NodeMap(size_t n_max) {
        size_t cap = 16;
        int bits = 4;
        while (cap < 2 * n_max) cap <<= 1, ++bits;
        key.assign(cap, 0xffffffffu);
        val.assign(cap, -1);
        mask = (uint32_t)(cap - 1);
        shift = 32 - bits;
      }
      // slot of g, inserted (val = -1) if absent; *fresh tells which
      uint32_t slot(uint32_t g, bool *fresh) {
        uint32_t i = (g * 2654435761u) >> shift;
        for (;; i = (i + 1) & mask) {
          if (key[i] == g) return *fresh = false, i;
          if (key[i] == 0xffffffffu) {
            key[i] = g;
            used.push_back(i);
            return *fresh = true, i;
          }
        }
      }
      void clear() {
        for (uint32_t i : used) key[i] = 0xffffffffu, val[i] = -1;
        used.clear();
      }
    };
    std::vector<std::vector<uint32_t>> pnode_of(T);       // per thread: node lists, padded
    std::vector<std::vector<int32_t>> shared_of(T);       // per thread: shared index per slot
    std::vector<int64_t> range_of(T + 1, n_patch);
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
      const int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
      const int tid = 0, nt = 1;
#endif
      const int64_t p0 = n_patch * tid / nt, p1 = n_patch * (tid + 1) / nt;
      range_of[tid] = p0;
      std::vector<uint32_t> &pn = pnode_of[tid];
      std::vector<int32_t> &so = shared_of[tid];
      NodeMap map((size_t)PE * NN);
      std::vector<uint32_t> priv, shar;
      bool fresh = false;
      for (int64_t p = p0; p < p1; ++p) {
        const int64_t s0 = p * PE, s1 = std::min<int64_t>(s0 + PE, n_order);
        priv.clear();
        shar.clear();
        for (int64_t s = s0; s < s1; ++s) {
          if (P->elem_of_slot[s] < 0) continue;
          const uint32_t *row = l2g + P->elem_of_slot[s] * NN;
          for (int k = 0; k < NN; ++k) {
            const uint32_t g = row[k];
            map.slot(g, &fresh);
            if (fresh) (multi[g] == 0 ? priv : shar).push_back(g);
          }
        }
        std::sort(priv.begin(), priv.end());
        std::sort(shar.begin(), shar.end());
        // node list order: [private | shared], each ascending; the patch writes the private
        // entries to the result vector itself
        const int32_t np = (int32_t)priv.size(), ns = (int32_t)shar.size();
        P->patch_npriv[p] = np;
        P->patch_nnodes[p] = np + ns;
        for (int32_t k = 0; k < np; ++k) {
          const uint32_t g = priv[k];
          map.val[map.slot(g, &fresh)] = k;
          pn.push_back((dirichlet && dirichlet[g]) ? (g | SEMK_NODE_DIRICHLET) : g);
        }
        for (int32_t k = 0; k < ns; ++k) {
          const uint32_t g = shar[k];
          map.val[map.slot(g, &fresh)] = np + k;
          uint32_t v = g | SEMK_NODE_SHARED;
          if (dirichlet && dirichlet[g]) v |= SEMK_NODE_DIRICHLET;
          pn.push_back(v);
          so.push_back(shared_index[g]);
        }
        // pad to a multiple of 4 entries: every patch's list starts 16-byte aligned (TMA)
        while (pn.size() & 3u) pn.push_back(0xffffffffu);
        // element-local index table (empty slots keep index 0: a valid node of the patch;
        // their geometric factors are zero, so they contribute exact zeros)
        uint16_t *eb = P->eloc.data() + (size_t)p * ES;  // this patch's [m][le][t] table
        for (int64_t s = s0; s < s1; ++s) {
          if (P->elem_of_slot[s] < 0) continue;
          const uint32_t *row = l2g + P->elem_of_slot[s] * NN;
          const int le = (int)(s - s0);
          for (int k = 0; k < NN; ++k) {
            const int m = k / n1, t = k - m * n1;
            eb[((size_t)m * PE + le) * n1 + t] = (uint16_t)map.val[map.slot(row[k], &fresh)];
          }
        }
        map.clear();
      }
    }
    int64_t max_patch_nodes = 0, n_slots = 0;
    {
      int64_t ptr = 0;
      for (int64_t p = 0; p < n_patch; ++p) {
        const int64_t nn = P->patch_nnodes[p], ns = nn - P->patch_npriv[p];
        P->patch_slot_base[p] = (int32_t)n_slots;
        n_slots += ns;
        if (n_slots > INT32_MAX) {
          delete P;
          semk_set_error("semk_hostplan_create: interface slot count overflows int32");
          return SEMK_ERR_UNSUPPORTED;
        }
        ptr += (nn + 3) & ~(int64_t)3;
        if (ptr > INT32_MAX) {
          delete P;
          semk_set_error("semk_hostplan_create: patch node table overflows int32");
          return SEMK_ERR_UNSUPPORTED;
        }
        P->patch_node_ptr[p + 1] = (int32_t)ptr;
        max_patch_nodes = std::max<int64_t>(max_patch_nodes, nn);
      }
      P->pnode.resize((size_t)ptr);
    }
    std::vector<int32_t> shared_of_slot(n_slots);  // shared index of the node behind every slot
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int t = 0; t < T; ++t) {
      if (range_of[t] >= n_patch) continue;
      const int64_t p0 = range_of[t];
      if (!pnode_of[t].empty())
        std::memcpy(P->pnode.data() + P->patch_node_ptr[p0], pnode_of[t].data(),
                    pnode_of[t].size() * sizeof(uint32_t));
      if (!shared_of[t].empty())
        std::memcpy(shared_of_slot.data() + P->patch_slot_base[p0], shared_of[t].data(),
                    shared_of[t].size() * sizeof(int32_t));
      std::vector<uint32_t>().swap(pnode_of[t]);
      std::vector<int32_t>().swap(shared_of[t]);
    }

    // CSR of interface slots per shared node (slots in ascending patch order = ascending
    // slot index)
    for (int64_t sl = 0; sl < n_slots; ++sl) P->shared_ptr[shared_of_slot[sl] + 1] += 1;
    for (int64_t i = 0; i < n_shared; ++i) P->shared_ptr[i + 1] += P->shared_ptr[i];
    P->shared_slot.assign(n_slots, 0);
    {
      std::vector<int32_t> fill(P->shared_ptr.begin(), P->shared_ptr.end() - 1);
      for (int64_t sl = 0; sl < n_slots; ++sl)
        P->shared_slot[fill[shared_of_slot[sl]]++] = (int32_t)sl;
    }
    std::vector<int32_t>().swap(shared_of_slot);

    // Device slots are PATCH-ordered: patch p writes the partial sums of its shared nodes
    // to the contiguous run [patch_slot_base[p], +n shared) (coalesced stores).
    //
    // Interface reduction tables.  Nodes shared by exactly two patches are grouped by
    // patch pair and cut into AFFINE CHUNKS of up to 32 nodes
    //     node_k = node0 + k*dn,  slotA_k = a0 + k*da,  slotB_k = b0 + k*db   (k < len)
    // (a whole patch edge of a structured mesh is one or a few chunks), 8 words each:
    //     {node0, dn, a0, da, b0, db, len, Dirichlet bit mask}.
    // One warp reduces one chunk with coalesced accesses and no per-node record.
    // Everything else (corner nodes touched by 3+ patches, isolated pairs) keeps a
    // per-node record {node id | flags, slot 0, slot 1, ext}; ext = 0xffffffff or an
    // offset into SHARED_EXT where {extra count, extra slots...} continue the list.
    // Slots are always summed in ascending patch order, so results are deterministic.
    {
      std::vector<int32_t> patch_of_slot(n_slots);
      for (int64_t p = 0; p < n_patch; ++p) {
        const int32_t s0 = P->patch_slot_base[p];
        const int32_t s1 = s0 + (P->patch_nnodes[p] - P->patch_npriv[p]);
        for (int32_t sidx = s0; sidx < s1; ++sidx) patch_of_slot[sidx] = (int32_t)p;
      }
      struct Pair {
        int32_t pa, pb, sa, sb;
        uint32_t node;
      };
      std::vector<Pair> pairs;
      std::vector<uint8_t> in_chunk(n_shared, 0);
      std::vector<int32_t> pair_index;  // shared index of each entry of `pairs`
      for (int64_t i = 0; i < n_shared; ++i) {
        const int32_t j0 = P->shared_ptr[i];
        if (P->shared_ptr[i + 1] - j0 != 2) continue;
        const int32_t sa = P->shared_slot[j0], sb = P->shared_slot[j0 + 1];
        pairs.push_back({patch_of_slot[sa], patch_of_slot[sb], sa, sb, P->shared_node[i]});
      }
      std::vector<int32_t> ord(pairs.size());
      for (size_t i = 0; i < ord.size(); ++i) ord[i] = (int32_t)i;
      // (a strict total order -- slot numbers are unique -- so any correct sort, parallel or
      // not, gives the same sequence)
      semk_plan_sort(ord.begin(), ord.end(), T, [&](int32_t x, int32_t y) {
        // higher patch first: the chunks that are complete once patches [0, P) have run
        // form a prefix of the table (staged host apply)
        const Pair &a = pairs[x], &b = pairs[y];
        if (a.pb != b.pb) return a.pb < b.pb;
        if (a.pa != b.pa) return a.pa < b.pa;
        return a.sa < b.sa;
      });
      std::vector<uint8_t> taken(pairs.size(), 0);
      size_t s = 0;
      while (s < ord.size()) {
        const Pair &f = pairs[ord[s]];
        size_t e = s + 1;
        int32_t dn = 0, da = 0, db = 0;
        if (e < ord.size()) {
          const Pair &g = pairs[ord[e]];
          if (g.pa == f.pa && g.pb == f.pb) {
            dn = (int32_t)((int64_t)(g.node & SEMK_NODE_ID_MASK) -
                           (int64_t)(f.node & SEMK_NODE_ID_MASK));
            da = g.sa - f.sa;
            db = g.sb - f.sb;
            ++e;
            while (e < ord.size() && e - s < 32) {
              const Pair &h = pairs[ord[e]], &q = pairs[ord[e - 1]];
              if (h.pa != f.pa || h.pb != f.pb) break;
              if ((int64_t)(h.node & SEMK_NODE_ID_MASK) - (int64_t)(q.node & SEMK_NODE_ID_MASK) !=
                      dn ||
                  h.sa - q.sa != da || h.sb - q.sb != db)
                break;
              ++e;
            }
          }
        }
        const size_t len = e - s;
        {
          uint32_t mask = 0;
          for (size_t k = 0; k < len; ++k) {
            if (pairs[ord[s + k]].node & SEMK_NODE_DIRICHLET) mask |= (1u << k);
            taken[ord[s + k]] = 1;
          }
          P->shared_chunk.push_back(f.node & SEMK_NODE_ID_MASK);
          P->shared_chunk.push_back((uint32_t)dn);
          P->shared_chunk.push_back((uint32_t)f.sa);
          P->shared_chunk.push_back((uint32_t)da);
          P->shared_chunk.push_back((uint32_t)f.sb);
          P->shared_chunk.push_back((uint32_t)db);
          P->shared_chunk.push_back((uint32_t)len);
          P->shared_chunk.push_back(mask);
          P->chunk_maxpatch.push_back(f.pb);
        }
        s = e;
      }
      // per-node records for everything not covered by a chunk, ordered by the highest
      // patch touching the node (then by id): {node | flags, count, slot 0..5}; counts above
      // 6 continue in SHARED_EXT at the offset stored in the last word ({extra slots...})
      std::vector<std::pair<int32_t, int64_t>> rec_order;  // (highest patch, shared index)
      size_t pi = 0;
      for (int64_t i = 0; i < n_shared; ++i) {
        const int32_t j0 = P->shared_ptr[i], cnt = P->shared_ptr[i + 1] - j0;
        if (cnt == 2) {
          const bool t = taken[pi++] != 0;
          if (t) continue;
        }
        rec_order.emplace_back(patch_of_slot[P->shared_slot[j0 + cnt - 1]], i);
      }
      std::sort(rec_order.begin(), rec_order.end());
      for (const auto &ro : rec_order) {
        const int64_t i = ro.second;
        const int32_t j0 = P->shared_ptr[i], cnt = P->shared_ptr[i + 1] - j0;
        uint32_t r[8] = {P->shared_node[i], (uint32_t)cnt, 0, 0, 0, 0, 0, 0};
        if (cnt <= 6) {
          for (int32_t j = 0; j < cnt; ++j) r[2 + j] = (uint32_t)P->shared_slot[j0 + j];
        } else {
          for (int32_t j = 0; j < 5; ++j) r[2 + j] = (uint32_t)P->shared_slot[j0 + j];
          r[7] = (uint32_t)P->shared_ext.size();
          for (int32_t j = j0 + 5; j < j0 + cnt; ++j)
            P->shared_ext.push_back((uint32_t)P->shared_slot[j]);
        }
        P->shared_rec.insert(P->shared_rec.end(), r, r + 8);
        P->rec_maxpatch.push_back(ro.first);
      }
      if (P->shared_ext.empty()) P->shared_ext.push_back(0);
      if (P->shared_rec.empty()) P->shared_rec.assign(8, 0xffffffffu);
      if (P->shared_chunk.empty()) P->shared_chunk.assign(8, 0);
    }

    // Uniform-stride device blocks, one TMA bulk copy each per patch:
    //   node block  = the patch's node list RELATIVE to its smallest node id (flags in
    //                 the top bits as in PNODE), 0xffffffff padded;
    //   index block = [m][le][t] patch-local indices (uint16).
    // Identical blocks are stored once (DEDUPLICATED): on a regularly numbered mesh every
    // interior patch has the same relative node list and the same index table, so the
    // kernel's table reads become L2 hits instead of DRAM traffic.  The per-patch part
    // is an 8-word header
    //   {n nodes, n private, first slot, 0, base node id, node block index, index block
    //    index, inverse block index}.
    const int64_t pn_stride = (max_patch_nodes + 3) & ~(int64_t)3;
    const int64_t el_stride = ((int64_t)NN * PE + 7) & ~(int64_t)7;
    // Inverse tables for the gather-style assembly: for every patch node the positions
    // of its element-local contributions inside the CTA's transposition scratch
    // (row m, thread le*n1 + t -> m*RS + le*n1 + t), ascending element slot, 0xffff
    // padded to INV_WIDTH = the largest number of elements of one patch meeting in a
    // node, rounded up to a multiple of 4 (4 on a structured mesh).
    const int RS = semk_scratch_row_stride(n1, PE);
    if ((int64_t)(n1 - 1) * RS + (int64_t)PE * n1 > 0xfffe) {
      delete P;
      semk_set_error("semk_hostplan_create: scratch positions overflow 16 bits");
      return SEMK_ERR_UNSUPPORTED;
    }
    int64_t inv_width = 4;
    {
      int64_t max_count = 0;  // largest number of element-local entries meeting in one patch node
#pragma omp parallel num_threads(T) reduction(max : max_count)
      {
        std::vector<int32_t> cnt(max_patch_nodes, 0);
#pragma omp for schedule(static)
        for (int64_t p = 0; p < n_patch; ++p) {
          const uint16_t *eb = P->eloc.data() + (size_t)p * ES;
          const int64_t live = std::min<int64_t>(PE, n_order - p * PE);
          std::fill(cnt.begin(), cnt.end(), 0);
          for (int m = 0; m < n1; ++m)
            for (int64_t le = 0; le < live; ++le)
              for (int t = 0; t < n1; ++t) {
                if (P->elem_of_slot[p * PE + le] < 0) continue;
                const int32_t c = ++cnt[eb[((size_t)m * PE + le) * n1 + t]];
                if (c > max_count) max_count = c;
              }
        }
      }
      if (max_count > inv_width) inv_width = (max_count + 3) & ~(int64_t)3;
    }
    const int64_t inv_stride = pn_stride * inv_width;  // uint16 entries per block
    P->patch_hdr.assign((size_t)n_patch * 8, 0u);
    {
      auto hash_bytes = [](const void *ptr, size_t n) {
        const unsigned char *b = static_cast<const unsigned char *>(ptr);
        uint64_t h = 1469598103934665603ull;  // FNV-1a
        for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
        return h;
      };
      struct Scratch {
        std::vector<uint32_t> pblk;
        std::vector<uint16_t> eblk, iblk;
        std::vector<int32_t> fill;
      };
      auto make_scratch = [&]() {
        Scratch w;
        w.pblk.resize(pn_stride);
        w.eblk.resize(el_stride);
        w.iblk.resize(inv_stride);
        w.fill.resize(max_patch_nodes);
        return w;
      };
      // the three device blocks of patch p (relative node list, index table, inverse table)
      auto gen_blocks = [&](int64_t p, Scratch &w, uint32_t *base_out, uint32_t *top_out) {
        const int32_t nn = P->patch_nnodes[p];
        const uint32_t *src = P->pnode.data() + P->patch_node_ptr[p];
        uint32_t base = SEMK_NODE_ID_MASK, top = 0;
        for (int32_t k = 0; k < nn; ++k) {
          base = std::min(base, src[k] & SEMK_NODE_ID_MASK);
          top = std::max(top, src[k] & SEMK_NODE_ID_MASK);
        }
        if (nn == 0) base = 0;
        std::fill(w.pblk.begin(), w.pblk.end(), 0xffffffffu);
        for (int32_t k = 0; k < nn; ++k)
          w.pblk[k] = ((src[k] & SEMK_NODE_ID_MASK) - base) | (src[k] & ~SEMK_NODE_ID_MASK);
        std::fill(w.eblk.begin(), w.eblk.end(), (uint16_t)0);
        std::copy(P->eloc.begin() + (size_t)p * ES, P->eloc.begin() + (size_t)p * ES + (size_t)NN * PE,
                  w.eblk.begin());
        std::fill(w.iblk.begin(), w.iblk.end(), (uint16_t)0xffff);
        std::fill(w.fill.begin(), w.fill.end(), 0);
        const int64_t live = std::min<int64_t>(PE, n_order - p * PE);
        for (int64_t le = 0; le < live; ++le) {    // ascending element slot: fixed sum order
          if (P->elem_of_slot[p * PE + le] < 0) continue;
          for (int m = 0; m < n1; ++m)
            for (int t = 0; t < n1; ++t) {
              const int32_t loc = w.eblk[((size_t)m * PE + le) * n1 + t];
              w.iblk[(size_t)loc * inv_width + w.fill[loc]++] = (uint16_t)(m * RS + le * n1 + t);
            }
        }
        if (base_out) *base_out = base;
        if (top_out) *top_out = top;
      };
      const size_t pn_bytes = (size_t)pn_stride * sizeof(uint32_t);
      const size_t el_bytes = (size_t)el_stride * sizeof(uint16_t);
      const size_t inv_bytes = (size_t)inv_stride * sizeof(uint16_t);
      P->patch_maxnode.assign(n_patch, 0);
      // Deduplication with exact comparison, in two levels.  Every thread takes a CONTIGUOUS
      // range of patches and keeps its own pool of distinct blocks, numbered in order of first
      // appearance (hash + memcmp against its candidates, as a single-threaded build would).
      // The pools are then merged in thread order = patch order, which numbers the global
      // blocks in order of first appearance over all patches -- the numbering of the
      // single-threaded build -- and the headers are renumbered.
      struct Pool {
        std::vector<uint32_t> pn;
        std::vector<uint16_t> el, inv;
        std::unordered_multimap<uint64_t, int32_t> pn_seen, el_seen, inv_seen;
        std::vector<uint64_t> pn_hash, el_hash, inv_hash;  // hash of every pooled block
      };
      auto find_or_add = [&](auto &seen_c, auto &pool, auto &pool_hash, const auto *blk,
                             int64_t stride, uint64_t hv) {
        const size_t bytes = (size_t)stride * sizeof(blk[0]);
        auto range = seen_c.equal_range(hv);
        for (auto it = range.first; it != range.second; ++it)
          if (std::memcmp(pool.data() + (size_t)it->second * stride, blk, bytes) == 0)
            return it->second;
        const int32_t idx = (int32_t)(pool.size() / (size_t)stride);
        pool.insert(pool.end(), blk, blk + stride);
        pool_hash.push_back(hv);
        seen_c.emplace(hv, idx);
        return idx;
      };
      std::vector<Pool> pools(T);
      std::vector<int64_t> range_begin(T + 1, n_patch);
#pragma omp parallel num_threads(T)
      {
#ifdef _OPENMP
        const int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
        const int tid = 0, nt = 1;
#endif
        const int64_t p0 = n_patch * tid / nt, p1 = n_patch * (tid + 1) / nt;
        range_begin[tid] = p0;
        Pool &Q = pools[tid];
        Scratch w = make_scratch();
        for (int64_t p = p0; p < p1; ++p) {
          uint32_t base = 0, top = 0;
          gen_blocks(p, w, &base, &top);
          P->patch_maxnode[p] = top;
          uint32_t *h = P->patch_hdr.data() + (size_t)p * 8;
          h[0] = (uint32_t)P->patch_nnodes[p];
          h[1] = (uint32_t)P->patch_npriv[p];
          h[2] = (uint32_t)P->patch_slot_base[p];
          h[4] = base;
          // (thread-local block numbers for now)
          h[5] = (uint32_t)find_or_add(Q.pn_seen, Q.pn, Q.pn_hash, w.pblk.data(), pn_stride,
                                       hash_bytes(w.pblk.data(), pn_bytes));
          h[6] = (uint32_t)find_or_add(Q.el_seen, Q.el, Q.el_hash, w.eblk.data(), el_stride,
                                       hash_bytes(w.eblk.data(), el_bytes));
          h[7] = (uint32_t)find_or_add(Q.inv_seen, Q.inv, Q.inv_hash, w.iblk.data(), inv_stride,
                                       hash_bytes(w.iblk.data(), inv_bytes));
        }
      }
      // merge: global number of every thread-local block
      std::vector<std::vector<int32_t>> map_pn(T), map_el(T), map_inv(T);
      {
        Pool G;  // only the `seen` maps and hash lists are used; the blocks go to P->...
        for (int t = 0; t < T; ++t) {
          const Pool &Q = pools[t];
          for (size_t i = 0; i < Q.pn_hash.size(); ++i)
            map_pn[t].push_back(find_or_add(G.pn_seen, P->pnblk, G.pn_hash,
                                            Q.pn.data() + i * (size_t)pn_stride, pn_stride,
                                            Q.pn_hash[i]));
          for (size_t i = 0; i < Q.el_hash.size(); ++i)
            map_el[t].push_back(find_or_add(G.el_seen, P->elblk, G.el_hash,
                                            Q.el.data() + i * (size_t)el_stride, el_stride,
                                            Q.el_hash[i]));
          for (size_t i = 0; i < Q.inv_hash.size(); ++i)
            map_inv[t].push_back(find_or_add(G.inv_seen, P->invblk, G.inv_hash,
                                             Q.inv.data() + i * (size_t)inv_stride, inv_stride,
                                             Q.inv_hash[i]));
        }
      }
#pragma omp parallel for num_threads(T) schedule(static)
      for (int64_t p = 0; p < n_patch; ++p) {
        int t = 0;
        while (t + 1 < T && p >= range_begin[t + 1]) ++t;
        uint32_t *h = P->patch_hdr.data() + (size_t)p * 8;
        h[5] = (uint32_t)map_pn[t][h[5]];
        h[6] = (uint32_t)map_el[t][h[6]];
        h[7] = (uint32_t)map_inv[t][h[7]];
      }
    }
    P->scalars[SEMK_PS_N_PN_UNIQUE] = (int64_t)(P->pnblk.size() / (size_t)pn_stride);
    P->scalars[SEMK_PS_N_EL_UNIQUE] = (int64_t)(P->elblk.size() / (size_t)el_stride);
    P->scalars[SEMK_PS_PN_STRIDE] = pn_stride;
    P->scalars[SEMK_PS_N_INV_UNIQUE] = (int64_t)(P->invblk.size() / (size_t)inv_stride);
    P->scalars[SEMK_PS_INV_WIDTH] = inv_width;
    P->scalars[SEMK_PS_INV_STRIDE] = inv_stride;
    {
      // counts of the interface tables (the vectors hold one dummy entry when empty)
      int64_t n_chunk = 0, n_rec = 0;
      if (!(P->shared_chunk.size() == 8 && P->shared_chunk[6] == 0))
        n_chunk = (int64_t)P->shared_chunk.size() / 8;
      if (!(P->shared_rec.size() == 8 && P->shared_rec[0] == 0xffffffffu))
        n_rec = (int64_t)P->shared_rec.size() / 8;
      P->scalars[SEMK_PS_N_SHARED_CHUNK] = n_chunk;
      P->scalars[SEMK_PS_N_SHARED_REC] = n_rec;
    }
    P->scalars[SEMK_PS_EL_STRIDE] = el_stride;

    P->scalars[SEMK_PS_N_PATCH] = n_patch;
    P->scalars[SEMK_PS_N_PNODE] = (int64_t)P->pnode.size();
    P->scalars[SEMK_PS_N_SLOTS] = n_slots;
    P->scalars[SEMK_PS_N_SHARED] = n_shared;
    P->scalars[SEMK_PS_MAX_PATCH_NODES] = max_patch_nodes;
    P->scalars[SEMK_PS_N_SLOT_ELEMS] = n_slot_elems;
    P->scalars[SEMK_PS_ELOC_STRIDE] = ES;
  } catch (const std::bad_alloc &) {
    delete P;
    semk_set_error("semk_hostplan_create: out of host memory");
    return SEMK_ERR_INVALID;
  }
  *out = P;
  return SEMK_OK;
}

extern "C" int64_t semk_hostplan_scalar(const semk_hostplan *plan, int which) {
  if (!plan || which < 0 || which >= SEMK_PS_COUNT) return -1;
  return plan->scalars[which];
}

template <class T>
static const void *vec_ptr(const std::vector<T> &v, int64_t *n_bytes) {
  if (n_bytes) *n_bytes = (int64_t)(v.size() * sizeof(T));
  return v.data();
}

extern "C" const void *semk_hostplan_array(const semk_hostplan *plan, int which,
                                           int64_t *n_bytes) {
  if (n_bytes) *n_bytes = 0;
  if (!plan) return nullptr;
  switch (which) {
    case SEMK_PA_PATCH_NODE_PTR: return vec_ptr(plan->patch_node_ptr, n_bytes);
    case SEMK_PA_PNODE: return vec_ptr(plan->pnode, n_bytes);
    case SEMK_PA_PATCH_NPRIV: return vec_ptr(plan->patch_npriv, n_bytes);
    case SEMK_PA_PATCH_SLOT_BASE: return vec_ptr(plan->patch_slot_base, n_bytes);
    case SEMK_PA_ELOC: return vec_ptr(plan->eloc, n_bytes);
    case SEMK_PA_ELEM_OF_SLOT: return vec_ptr(plan->elem_of_slot, n_bytes);
    case SEMK_PA_SHARED_NODE: return vec_ptr(plan->shared_node, n_bytes);
    case SEMK_PA_SHARED_PTR: return vec_ptr(plan->shared_ptr, n_bytes);
    case SEMK_PA_SHARED_SLOT: return vec_ptr(plan->shared_slot, n_bytes);
    case SEMK_PA_PATCH_NNODES: return vec_ptr(plan->patch_nnodes, n_bytes);
    case SEMK_PA_PNBLK: return vec_ptr(plan->pnblk, n_bytes);
    case SEMK_PA_ELBLK: return vec_ptr(plan->elblk, n_bytes);
    case SEMK_PA_SHARED_REC: return vec_ptr(plan->shared_rec, n_bytes);
    case SEMK_PA_SHARED_EXT: return vec_ptr(plan->shared_ext, n_bytes);
    case SEMK_PA_SHARED_CHUNK: return vec_ptr(plan->shared_chunk, n_bytes);
    case SEMK_PA_PATCH_HDR: return vec_ptr(plan->patch_hdr, n_bytes);
    case SEMK_PA_INVBLK: return vec_ptr(plan->invblk, n_bytes);
    case SEMK_PA_PATCH_MAXNODE: return vec_ptr(plan->patch_maxnode, n_bytes);
    case SEMK_PA_CHUNK_MAXPATCH: return vec_ptr(plan->chunk_maxpatch, n_bytes);
    case SEMK_PA_REC_MAXPATCH: return vec_ptr(plan->rec_maxpatch, n_bytes);
    default: return nullptr;
  }
}

extern "C" void semk_hostplan_destroy(semk_hostplan *plan) { delete plan; }
