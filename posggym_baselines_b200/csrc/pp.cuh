// PredatorPrey-v0 (2 predators, 3 prey) on packed integer state.  Semantics follow
// posggym's PredatorPreyModel.step as consumed at planning/mcts.py:333,418 and
// planning/belief.py:165 of the reference.
//   state word 0: predator 0 (x | y<<4) | predator 1 << 8 | caught bits << 16
//   state word 1: prey 0 (x | y<<4) | prey 1 << 8 | prey 2 << 16
//   observation : 25 cells x 2 bits (EMPTY=0 WALL=1 PREDATOR=2 PREY=3), cell index
//                 row*5 + col of the 5x5 window centred on the predator (row = y offset; the order the shipped
//                 best-response policies were trained on: oracle/predator_prey.py, tests/test_oracle_policies.py)
#pragma once
#include <stdint.h>

#include "philox.cuh"

namespace posg {
namespace pp {

constexpr int PAD = 4, LINES = 24;
constexpr int NP = 2, NQ = 3;
enum : int { EMPTY = 0, WALL = 1, PREDATOR = 2, PREY = 3 };

struct alignas(16) Blob {
  uint32_t rows[LINES];            // rows[y+PAD] bit (x+PAD): blocked or outside
  uint8_t pred_start[16], prey_start[32];
  uint8_t n_pred_start, n_prey_start, size, obs_dim, prey_strength, cooperative, pad_[10];
  double reward_k[4];              // k * (1 / num_prey), k newly caught prey (cooperative)
  double per_prey_reward;
  double pad2_;
  double share_k[4];               // per_prey_reward / k, k predators involved in a capture (non-cooperative)
  unsigned long long obs_base[256];  // wall cells of the window centred at (x, y)
  uint8_t wall4[256];                // bit k: neighbour k (N, E, S, W) of cell x | y << 4 is blocked / outside
  uint8_t nth_lut[32 * 8];           // [mask5 * 8 + n] -> index of the n-th set bit of mask5 (bits 0..4)
  int8_t nb_delta[8];                // cell-index delta of N, E, S, W, stay
  uint8_t lut_ok, pad3_[7];          // 1: the relative-offset tables below are exact for this grid (size <= 10)
  // ---- relative-offset tables, index REL0 + (b - a) for packed cells a, b (x | y << 4).  The offset
  // dx + 16 dy is unique for |dx| <= 2 on grids up to 14 wide; caught prey are given the cell SENT,
  // which on grids up to 10x10 is "far" from every cell under every table (no select per use).
  uint8_t rel_nb[512];               // bit k: b is the N / E / S / W neighbour of a
  // threat score of b for a prey at a: 4 * Manhattan + 2 within obs_dim, else THR_FAR (the caller adds +-1 to
  // the second of two candidates -- the tie pick -- and THR_PREY to prey candidates, then takes a minimum
  // together with the score THR_RND of "nobody in sight: random move")
  uint8_t rel_thr[512];
  // cell delta of a prey's move, [row * 16 + free-neighbour mask (N, E, S, W)]: rows MOVE_FLEE0 + d = flee
  // (_move_away) from a threat at offset d; rows MOVE_RND + picks8 = random move with picks8 = p2 | p3 << 1 |
  // p4 << 3 | p5 << 5, p_n = pick_index(word, n)
  int8_t move_delta[(80 + 256) * 16];
  uint8_t pred_move[5 * 256];        // [action * 256 + cell] -> the predator's target cell (walls resolved)
  uint32_t caught_mask[8];           // bytes 0xFF at the prey whose caught bit is set
  // observation pieces: window cells that are outside the grid by column (+ the centre PREDATOR) / by row,
  // and the PREY code at the window cell of an entity at offset u = (dx + od) + 16 (dy + od), u < 80
  unsigned long long win_col[16], win_row[16];
  unsigned long long win_pos[256];   // win_col[p & 15] | win_row[p >> 4] per packed cell p: one lookup per observation
  unsigned long long win_obs[128];   // index (dx + 2) + 16 (dy + 2) < 128 without bit 3; zero outside the 5 x 5 window
};
constexpr int REL0 = 256, SENT = 0xCC, THR_FAR = 66, THR_PREY = 32, THR_RND = 64;
// move_delta rows are addressed in the biased offset space REL0 + d of the candidate keys: flee rows sit at
// REL0 + d - MOVE_BASE (|d| <= 34), the random rows behind them
constexpr int MOVE_BASE = REL0 - 40, MOVE_RND = 80;

// host side: fill the lookup tables from rows[] (walls)
inline void build_tables(Blob* b) {
  for (int y = 0; y < 16; ++y)
    for (int x = 0; x < 16; ++x) {
      const int nx[4] = {x, x + 1, x, x - 1}, ny[4] = {y - 1, y, y + 1, y};
      uint8_t m = 0;
      for (int k = 0; k < 4; ++k) m |= (uint8_t)(((b->rows[ny[k] + PAD] >> (nx[k] + PAD)) & 1u) << k);
      b->wall4[x | (y << 4)] = m;
    }
  for (int m = 0; m < 32; ++m) {
    int n = 0;
    for (int bit = 0; bit < 5; ++bit) if ((m >> bit) & 1) b->nth_lut[m * 8 + n++] = (uint8_t)bit;
    for (; n < 8; ++n) b->nth_lut[m * 8 + n] = 4;
  }
  const int8_t d[8] = {-16, 1, 16, -1, 0, 0, 0, 0};
  for (int k = 0; k < 8; ++k) b->nb_delta[k] = d[k];
  for (int k = 0; k < 4; ++k) b->share_k[k] = b->per_prey_reward / (double)(k < 1 ? 1 : k);
  // relative-offset tables
  const int od = b->obs_dim, w = 2 * od + 1;
  bool blocks = false;
  for (int y = 0; y < b->size; ++y)
    for (int x = 0; x < b->size; ++x) blocks |= ((b->rows[y + PAD] >> (x + PAD)) & 1u) != 0;
  b->lut_ok = (b->size <= 10 && od == 2 && !blocks) ? 1 : 0;
  for (int i = 0; i < 512; ++i) { b->rel_nb[i] = 0; b->rel_thr[i] = (uint8_t)THR_FAR; }
  for (int i = 0; i < 128; ++i) b->win_obs[i] = 0;
  for (int dy = -od; dy <= od; ++dy)
    for (int dx = -od; dx <= od; ++dx) {
      const int i = REL0 + dx + 16 * dy, man = abs(dx) + abs(dy);
      if (man <= od && man > 0) b->rel_thr[i] = (uint8_t)(4 * man + 2);
      if (man == 1) b->rel_nb[i] = (uint8_t)(dy == -1 ? 1 : dx == 1 ? 2 : dy == 1 ? 4 : 8);
      b->win_obs[(dx + od) + 16 * (dy + od)] = (unsigned long long)PREY << (2 * ((dy + od) * w + (dx + od)));
    }
  for (int v = 0; v < 16; ++v) {
    unsigned long long col = (unsigned long long)PREDATOR << (2 * (od * w + od)), row = 0;
    for (int k = 0; k < w; ++k) {
      const int g = v + k - od;   // grid column (row) shown in window column (row) k
      if (g >= 0 && g < b->size) continue;
      for (int m = 0; m < w; ++m) {
        col |= (unsigned long long)WALL << (2 * (m * w + k));   // window column k, every row m
        row |= (unsigned long long)WALL << (2 * (k * w + m));   // window row k, every column m
      }
    }
    b->win_col[v] = col; b->win_row[v] = row;
  }
  for (int p = 0; p < 256; ++p) b->win_pos[p] = b->win_col[p & 15] | b->win_row[p >> 4];
  for (int c = 0; c < 8; ++c)
    b->caught_mask[c] = ((c & 1) ? 0xFFu : 0u) | ((c & 2) ? 0xFF00u : 0u) | ((c & 4) ? 0xFF0000u : 0u);
  // prey moves.  Flee: keep the free option (staying included) with the largest (distance to the threat, x, y)
  const int ox[5] = {0, 1, 0, -1, 0}, oy[5] = {-1, 0, 1, 0, 0};
  for (int i = 0; i < (80 + 256) * 16; ++i) b->move_delta[i] = 0;
  for (int dy = -od; dy <= od; ++dy)
    for (int dx = -od; dx <= od; ++dx) {
      if (abs(dx) + abs(dy) > od) continue;
      for (int f = 0; f < 16; ++f) {
        int best = 4, bd = abs(dx) + abs(dy);
        for (int k = 0; k < 4; ++k) {
          if (!((f >> k) & 1)) continue;
          const int dd = abs(dx - ox[k]) + abs(dy - oy[k]);
          const bool better = dd > bd || (dd == bd && (ox[k] > ox[best] || (ox[k] == ox[best] && oy[k] > oy[best])));
          if (better) { best = k; bd = dd; }
        }
        b->move_delta[(REL0 + dx + 16 * dy - MOVE_BASE) * 16 + f] = d[best];
      }
    }
  // random move: the pick-th of the free neighbours N, E, S, W followed by staying
  for (int picks = 0; picks < 256; ++picks)
    for (int f = 0; f < 16; ++f) {
      const int pk[6] = {0, 0, picks & 1, (picks >> 1) & 3, (picks >> 3) & 3, (picks >> 5) & 7};
      const int m5 = f | 16;
      int cnt = 0;
      for (int bit = 0; bit < 5; ++bit) cnt += (m5 >> bit) & 1;
      b->move_delta[(MOVE_RND + picks) * 16 + f] = d[b->nth_lut[m5 * 8 + (pk[cnt] < cnt ? pk[cnt] : cnt - 1)]];
    }
  // predator targets
  for (int a = 0; a < 5; ++a)
    for (int c = 0; c < 256; ++c) {
      const int dir = a == 1 ? 0 : a == 2 ? 2 : a == 3 ? 3 : a == 4 ? 1 : 4;   // UP, DOWN, LEFT, RIGHT -> N, S, W, E
      const bool blocked = dir < 4 && ((b->wall4[c] >> dir) & 1);
      b->pred_move[a * 256 + c] = (uint8_t)(c + ((dir == 4 || blocked) ? 0 : (int)d[dir]));
    }
}
static_assert(sizeof(Blob) % 16 == 0, "blob must be a multiple of 16 B");

__device__ __forceinline__ bool blocked(const Blob& B, int x, int y) {
  return (B.rows[y + PAD] >> (x + PAD)) & 1u;
}
__device__ __forceinline__ int mdist(int a, int b) {
  return abs((a & 15) - (b & 15)) + abs((a >> 4) - (b >> 4));
}
__device__ __forceinline__ uint32_t pick_index(uint32_t word, uint32_t n) { return __umulhi(word, n); }

struct Ents {           // unpacked entities
  int pred[NP];
  int prey[NQ];
  uint32_t caught;      // bit i
};

__device__ __forceinline__ void unpack(uint32_t w0, uint32_t w1, Ents& e) {
  e.pred[0] = w0 & 255; e.pred[1] = (w0 >> 8) & 255; e.caught = (w0 >> 16) & 7u;
  e.prey[0] = w1 & 255; e.prey[1] = (w1 >> 8) & 255; e.prey[2] = (w1 >> 16) & 255;
}
__device__ __forceinline__ void pack(const Ents& e, uint32_t& w0, uint32_t& w1) {
  w0 = (uint32_t)e.pred[0] | ((uint32_t)e.pred[1] << 8) | (e.caught << 16);
  w1 = (uint32_t)e.prey[0] | ((uint32_t)e.prey[1] << 8) | ((uint32_t)e.prey[2] << 16);
}

// |ax - bx| + |ay - by| in two instructions (VABSDIFF with accumulate) on unpacked coordinates
__device__ __forceinline__ int mdist_xy(int ax, int ay, int bx, int by) {
  return (int)__sad(ax, bx, __sad(ay, by, 0u));
}

// One prey move (PredatorPreyModel._move_away / the random move of _get_next_prey_state), branch
// free: flee = keep the free option (staying included) with the largest (distance to `threat`, x, y);
// otherwise the pick_index(word)-th of the free neighbours N, E, S, W followed by staying.
// occ[] = cells of the predators and of the uncaught prey (moves made so far applied; -1 = none).
// The (x, y) tie-break among equally distant options is a fixed priority of the five options:
// E (x+1) > S (y+1) > stay > N (y-1) > W (x-1).
__device__ __forceinline__ int prey_move(const Blob& B, int coord, bool flee, int threat, uint32_t word,
                                         const int (&occ)[NP + NQ]) {
  const uint32_t w4 = B.wall4[coord];
  const int dx = (threat & 15) - (coord & 15), dy = (threat >> 4) - (coord >> 4);
  const int adx = abs(dx), ady = abs(dy);
  const int nb[4] = {coord - 16, coord + 1, coord + 16, coord - 1};
  const int dist[4] = {adx + abs(dy + 1), abs(dx - 1) + ady, adx + abs(dy - 1), abs(dx + 1) + ady};
  constexpr uint32_t PRIO[4] = {1u, 4u, 3u, 0u};       // N, E, S, W; staying = 2
  uint32_t free4 = 0, best = ((uint32_t)(adx + ady) << 3) | 2u;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    bool taken = false;
#pragma unroll
    for (int j = 0; j < NP + NQ; ++j) taken |= (nb[k] == occ[j]);
    const bool fr = !((w4 >> k) & 1u) && !taken;
    free4 |= (fr ? 1u : 0u) << k;
    best = max(best, fr ? (((uint32_t)dist[k] << 3) | PRIO[k]) : 0u);
  }
  // priority -> option index (W, N, stay, S, E -> 3, 0, 4, 2, 1), packed three bits each
  const uint32_t flee_sel = (0x1503u >> (3u * (best & 7u))) & 7u;
  const uint32_t m5 = free4 | 16u;
  const uint32_t rnd_sel = B.nth_lut[m5 * 8u + pick_index(word, (uint32_t)__popc(m5))];
  return coord + (int)B.nb_delta[flee ? flee_sel : rnd_sel];
}

// joint transition; rew[] per predator; returns all_done
__device__ __forceinline__ bool transition(const Blob& B, Ents& e, const int (&act)[NP], const uint32_t (&words)[NQ],
                                           double (&rew)[NP]) {
  const int od = (int)B.obs_dim;
  const int n_unc = NQ - __popc(e.caught);
  // Which rule moves each prey is a function of the state before the step: 0 flee from the closest
  // predator in sight, 1 keep away from the closest other prey in sight, 2 random move, 3 caught.
  // The moves are then made in (rule, index) order, as the reference's three passes do.
  uint32_t info[NQ], key[NQ];   // info: cell | threat << 8 | flee << 16
  const int px[NP] = {e.pred[0] & 15, e.pred[1] & 15}, py[NP] = {e.pred[0] >> 4, e.pred[1] >> 4};
  const int qx[NQ] = {e.prey[0] & 15, e.prey[1] & 15, e.prey[2] & 15};
  const int qy[NQ] = {e.prey[0] >> 4, e.prey[1] >> 4, e.prey[2] >> 4};
  // the three prey-prey distances, shared by both ends
  const int qq01 = mdist_xy(qx[0], qy[0], qx[1], qy[1]), qq02 = mdist_xy(qx[0], qy[0], qx[2], qy[2]);
  const int qq12 = mdist_xy(qx[1], qy[1], qx[2], qy[2]);
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int c = e.prey[i];
    const bool caught = (e.caught >> i) & 1u;
    const bool hi = (words[i] >> 31) != 0u;            // pick_index(word, 2)
    const int d0 = mdist_xy(qx[i], qy[i], px[0], py[0]), d1 = mdist_xy(qx[i], qy[i], px[1], py[1]);
    const bool ph1 = min(d0, d1) <= od;
    const int t1 = d0 < d1 ? e.pred[0] : (d1 < d0 ? e.pred[1] : (hi ? e.pred[1] : e.pred[0]));
    const int j1 = i == 0 ? 1 : 0, j2 = i == 2 ? 1 : 2;
    const int e1 = i == 0 ? qq01 : (i == 1 ? qq01 : qq02), e2 = i == 0 ? qq02 : (i == 1 ? qq12 : qq12);
    const int f1 = ((e.caught >> j1) & 1u) ? (1 << 20) : e1;
    const int f2 = ((e.caught >> j2) & 1u) ? (1 << 20) : e2;
    const bool ph2 = !ph1 && n_unc > 1 && min(f1, f2) <= od;
    const int t2 = f1 < f2 ? e.prey[j1] : (f2 < f1 ? e.prey[j2] : (hi ? e.prey[j2] : e.prey[j1]));
    const uint32_t rule = caught ? 3u : (ph1 ? 0u : (ph2 ? 1u : 2u));
    info[i] = (uint32_t)c | ((uint32_t)(ph1 ? t1 : t2) << 8) | ((ph1 || ph2) ? 0x10000u : 0u);
    key[i] = rule * 4u + (uint32_t)i;
  }
  // sort the three keys
  uint32_t k0 = min(key[0], key[1]), k1 = max(key[0], key[1]), k2 = key[2];
  { const uint32_t t = min(k1, k2); k2 = max(k1, k2); k1 = t; }
  { const uint32_t t = min(k0, k1); k1 = max(k0, k1); k0 = t; }
  const uint32_t order[NQ] = {k0, k1, k2};
  int cur[NQ] = {e.prey[0], e.prey[1], e.prey[2]};
  int occ[NP + NQ] = {e.pred[0], e.pred[1], (e.caught & 1u) ? -1 : cur[0], (e.caught & 2u) ? -1 : cur[1],
                      (e.caught & 4u) ? -1 : cur[2]};
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    const uint32_t i = order[k] & 3u;
    const bool live = order[k] < 12u;
    const uint32_t inf = i == 0u ? info[0] : (i == 1u ? info[1] : info[2]);
    const uint32_t word = i == 0u ? words[0] : (i == 1u ? words[1] : words[2]);
    const int nxt = prey_move(B, (int)(inf & 255u), (inf >> 16) != 0u, (int)((inf >> 8) & 255u), word, occ);
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      const bool mine = live && i == (uint32_t)j;
      cur[j] = mine ? nxt : cur[j];
      occ[NP + j] = mine ? nxt : occ[NP + j];
    }
  }
  // predators: simultaneous moves; prey cells and shared targets block
  int pot[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int a = act[i];
    const uint32_t dir = (0x13200u >> (4 * a)) & 3u;   // UP, DOWN, LEFT, RIGHT -> N, S, W, E
    const bool moves = a != 0 && !((B.wall4[e.pred[i]] >> dir) & 1u);
    const int c = moves ? e.pred[i] + (int)B.nb_delta[dir] : e.pred[i];
    const bool on_prey = (c == occ[NP]) || (c == occ[NP + 1]) || (c == occ[NP + 2]);
    pot[i] = on_prey ? e.pred[i] : c;
  }
  const bool clash = pot[0] == pot[1];
  const int p0 = clash ? e.pred[0] : pot[0], p1 = clash ? e.pred[1] : pot[1];
  // captures
  uint32_t newly = 0;
  const int p0x = p0 & 15, p0y = p0 >> 4, p1x = p1 & 15, p1y = p1 >> 4;
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    const int cx = cur[j] & 15, cy = cur[j] >> 4;
    const int adj = (mdist_xy(cx, cy, p0x, p0y) <= 1) + (mdist_xy(cx, cy, p1x, p1y) <= 1);
    newly |= (!((e.caught >> j) & 1u) && adj >= (int)B.prey_strength) ? (1u << j) : 0u;
  }
  rew[0] = 0.0; rew[1] = 0.0;
  if (B.cooperative) {
    rew[0] = rew[1] = B.reward_k[__popc(newly)];   // reward_k[0] == 0.0
  } else if (newly) {
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      if (!((newly >> j) & 1u)) continue;
      const bool i0 = mdist(cur[j], p0) <= 1, i1 = mdist(cur[j], p1) <= 1;
      const double share = __ddiv_rn(B.per_prey_reward, (double)((int)i0 + (int)i1));
      if (i0) rew[0] = __dadd_rn(rew[0], share);
      if (i1) rew[1] = __dadd_rn(rew[1], share);
    }
  }
  e.pred[0] = p0; e.pred[1] = p1;
#pragma unroll
  for (int j = 0; j < NQ; ++j) e.prey[j] = cur[j];
  e.caught |= newly;
  return e.caught == 7u;
}

// window cell index (row*5 + col) of grid cell c seen from predator at p, or -1
__device__ __forceinline__ int window_cell(int p, int c, int obs_dim) {
  const int col = (c & 15) - (p & 15) + obs_dim, row = (c >> 4) - (p >> 4) + obs_dim;
  const int w = 2 * obs_dim + 1;
  return ((unsigned)col < (unsigned)w && (unsigned)row < (unsigned)w) ? row * w + col : -1;
}

__device__ __forceinline__ unsigned long long observe(const Blob& B, const Ents& e, int i) {
  const int p = e.pred[i];
  unsigned long long cells = B.obs_base[p];
  const int od = B.obs_dim;
  // Entities stand on free cells (EMPTY = 0 in the wall-only base) and never share a cell -- moves
  // into occupied cells are rejected by the transition -- so each one is a single OR.
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    const int k = ((e.caught >> j) & 1u) ? -1 : window_cell(p, e.prey[j], od);
    cells |= k >= 0 ? ((unsigned long long)PREY << (2 * k)) : 0ULL;
  }
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const int k = window_cell(p, e.pred[j], od);
    cells |= k >= 0 ? ((unsigned long long)PREDATOR << (2 * k)) : 0ULL;
  }
  return cells;
}


// i == 0 ? a : (i == 1 ? b : c) as two selects (left to itself the compiler branches around them)
__device__ __forceinline__ uint32_t sel3(uint32_t i, uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("{\n\t.reg .pred p0, p1;\n\tsetp.eq.u32 p0, %1, 0;\n\tsetp.eq.u32 p1, %1, 1;\n\t"
      "selp.b32 %0, %3, %4, p1;\n\tselp.b32 %0, %2, %0, p0;\n\t}" : "=r"(r) : "r"(i), "r"(a), "r"(b), "r"(c));
  return r;
}

// ---- table-driven transition (block-free grids up to 10x10: Blob::lut_ok) ---------------------
// Same semantics as transition() above, on the packed words, balanced between the two units that bound
// these kernels: the integer ALU pipe and the shared-memory data pipe (a random byte lookup costs ~2-5
// wavefronts of bank conflicts = as much as 4-10 ALU instructions, so tables are used only where they
// replace more than that):
//  * neighbour bits / threat scores are byte lookups on the packed-cell difference; caught prey sit on the
//    cell SENT, which is far from everything, has no free neighbour and is adjacent to nothing, so nothing
//    below needs a "caught" select (a caught prey "moves" by 0 and is "captured" never);
//  * a prey's rule, threat and move-table row are ONE minimum over five candidate keys score << 10 | row
//    (closest predator, else closest other prey -- the tie pick folded into the score as +-1 --, else the
//    random-move row holding its pick for every possible number of options);
//  * the sort key of a prey carries rule class, index, cell and that row, so the move is
//    move_delta[row * 16 + free-neighbour mask] with no dynamic select;
//  * adjacency for the captures is arithmetic: the packed difference d of two cells on a grid this small
//    has d * d in {0, 1, 256} exactly when they are within Manhattan distance 1.
// occ[] returns the cells of the prey after the step with SENT for every caught one (observe_lut).
__device__ __forceinline__ bool transition_lut(const Blob& B, uint32_t& w0, uint32_t& w1, const int (&act)[NP],
                                               const uint32_t (&words)[NQ], double (&rew)[NP], int (&occ)[NQ]) {
  const uint8_t* __restrict__ THR = B.rel_thr + REL0;
  const uint8_t* __restrict__ NB = B.rel_nb + REL0;
  const int p0 = (int)(w0 & 255u), p1 = (int)((w0 >> 8) & 255u);
  const uint32_t caught = (w0 >> 16) & 7u;
  const uint32_t cmask = B.caught_mask[caught];
  const uint32_t cells = (w1 & ~cmask) | ((uint32_t)SENT * 0x010101u & cmask);
  occ[0] = (int)__byte_perm(cells, 0u, 0x4440); occ[1] = (int)__byte_perm(cells, 0u, 0x4441);
  occ[2] = (int)__byte_perm(cells, 0u, 0x4442);
  // prey-prey threat scores, shared by both ends
  const int t01 = THR[occ[1] - occ[0]], t02 = THR[occ[2] - occ[0]], t12 = THR[occ[2] - occ[1]];
  uint32_t key[NQ];   // (class * 4 + index) << 24 | cell << 16 | move-table row: moves are made in ascending order
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int c = occ[i], rc = REL0 - c;
    const int sgn = (int)words[i] >> 31;                 // -pick_index(word, 2): the tie pick
    const int j1 = i == 0 ? 1 : 0, j2 = i == 2 ? 1 : 2;
    const int f1 = i == 2 ? t02 : t01, f2 = i == 0 ? t02 : t12;
    // candidate keys score << 10 | biased offset: the second of a pair wins a tie iff the pick says so
    // (score - 1 instead of + 1)
    const int second = (sgn & (int)0xFFFFF800u) | 0x400;   // 2048 * sgn + 1024 = +-1024 in one LOP3
    const int d0 = p0 + rc, d1 = p1 + rc, e1 = occ[j1] + rc, e2 = occ[j2] + rc;
    const int k0 = THR[d0 - REL0] * 1024 + d0;
    const int k1 = THR[d1 - REL0] * 1024 + (second + d1);
    const int k2 = f1 * 1024 + (THR_PREY * 1024 + e1);
    const int k3 = f2 * 1024 + (THR_PREY * 1024 + second + e2);
    // pick_index(word, n) for n = 2..5 options, packed: the random-move row
    const int k4 = (((int)__umulhi(words[i], 5u) * 4 + (int)(words[i] >> 30)) * 4 + (int)__umulhi(words[i], 3u)) * 2 +
                   (THR_RND * 1024 + MOVE_BASE + MOVE_RND - sgn);
    const int best = min(min(min(k0, k1), min(k2, k3)), k4);
    // class = score >> 5: 0 predator in sight, 1 prey in sight, 2 random
    key[i] = (uint32_t)((best >> 15) * (4 << 24) + (c * 65536 + ((best & 1023) + (i << 24))));
  }
  uint32_t k0 = min(key[0], key[1]), k1 = max(key[0], key[1]), k2 = key[2];
  { const uint32_t t = min(k1, k2); k2 = max(k1, k2); k1 = t; }
  { const uint32_t t = min(k0, k1); k1 = max(k0, k1); k0 = t; }
  const uint32_t order[NQ] = {k0, k1, k2};
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    const int c = (int)__byte_perm(order[k], 0u, 0x4442);
    const int row = (int)(order[k] & 0xFFFFu);
    // a prey is no neighbour of itself (offset 0), so all five entities go through the table
    const uint32_t taken = NB[p0 - c] | NB[p1 - c] | NB[occ[0] - c] | NB[occ[1] - c] | NB[occ[2] - c];
    const uint32_t free4 = ~((uint32_t)B.wall4[c] | taken) & 15u;
    const int nxt = c + (int)B.move_delta[(row - MOVE_BASE) * 16 + (int)free4];
#pragma unroll
    for (int j = 0; j < NQ; ++j) occ[j] = (((order[k] >> 24) & 3u) == (uint32_t)j) ? nxt : occ[j];
  }
  // predators: simultaneous moves; prey cells and shared targets block
  int pot[NP];
  const int pc[NP] = {p0, p1};
  // (the prey cells as three bytes: "c equals one of them" is the zero-byte test on the XOR with c in every byte)
  const uint32_t occ3 = (uint32_t)occ[0] + ((uint32_t)occ[1] * 256u + (uint32_t)occ[2] * 65536u);
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int c = (int)B.pred_move[act[i] * 256 + pc[i]];
    const uint32_t x = occ3 ^ ((uint32_t)c * 0x010101u);
    const bool on_prey = ((x - 0x010101u) & ~x & 0x808080u) != 0u;
    pot[i] = on_prey ? pc[i] : c;
  }
  const bool clash = pot[0] == pot[1];
  const int n0 = clash ? p0 : pot[0], n1 = clash ? p1 : pot[1];
  // captures (SENT is adjacent to nothing): within Manhattan distance 1 <=> d * d in {0, 1, 256}
  uint32_t sq[NQ][NP];
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    const int da = n0 - occ[j], db = n1 - occ[j];
    sq[j][0] = (uint32_t)(da * da); sq[j][1] = (uint32_t)(db * db);
  }
  uint32_t newly;
  if (B.prey_strength >= 2)   // uniform: both predators adjacent
    newly = ((((sq[0][0] | sq[0][1]) & 0xFFFFFEFEu) == 0u) ? 1u : 0u) | ((((sq[1][0] | sq[1][1]) & 0xFFFFFEFEu) == 0u) ? 2u : 0u) |
            ((((sq[2][0] | sq[2][1]) & 0xFFFFFEFEu) == 0u) ? 4u : 0u);
  else                        // a product with a zero factor: one of them adjacent
    newly = (((sq[0][0] & 0xFFFFFEFEu) * (sq[0][1] & 0xFFFFFEFEu) == 0u) ? 1u : 0u) |
            (((sq[1][0] & 0xFFFFFEFEu) * (sq[1][1] & 0xFFFFFEFEu) == 0u) ? 2u : 0u) |
            (((sq[2][0] & 0xFFFFFEFEu) * (sq[2][1] & 0xFFFFFEFEu) == 0u) ? 4u : 0u);
  rew[0] = 0.0; rew[1] = 0.0;
  if (B.cooperative) {
    rew[0] = rew[1] = B.reward_k[__popc(newly)];   // reward_k[0] == 0.0
  } else if (newly) {
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      if (!((newly >> j) & 1u)) continue;
      const int da = n0 - occ[j], db = n1 - occ[j];
      const bool i0 = ((uint32_t)(da * da) & 0xFFFFFEFEu) == 0u, i1 = ((uint32_t)(db * db) & 0xFFFFFEFEu) == 0u;
      const double share = B.share_k[(int)i0 + (int)i1];
      if (i0) rew[0] = __dadd_rn(rew[0], share);
      if (i1) rew[1] = __dadd_rn(rew[1], share);
    }
  }
  // caught prey keep the cell they had; newly caught ones the cell they were caught on
  const uint32_t moved = occ3;
  const uint32_t all = caught | newly;
  w0 = (uint32_t)n0 + ((uint32_t)n1 * 256u + all * 65536u);
  w1 = (w1 & cmask) | (moved & ~cmask);
#pragma unroll
  for (int j = 0; j < NQ; ++j) occ[j] = ((newly >> j) & 1u) ? SENT : occ[j];
  return all == 7u;
}

// observation of predator i: out-of-grid columns / rows (and the own cell) from one table indexed by the packed
// cell; another entity contributes the table entry of its offset, looked up only by the lanes that have it in
// (an 8 x 8 superset of) the window: a predicated-off lane makes no shared-memory access, so the lookup costs a
// wavefront or two, not six
__device__ __forceinline__ unsigned long long observe_lut(const Blob& B, uint32_t w0, const int (&occ)[NQ], int i) {
  const uint32_t p = (w0 >> (8 * i)) & 255u, o = (w0 >> (8 * (1 - i))) & 255u;
  const uint32_t bias = 2u * 17u - p;     // obs_dim = 2: u = (dx + 2) + 16 (dy + 2)
  // every entity's contribution is a value (zero when it is outside the window: the lookup is predicated off), so
  // the six words OR together in three-input LOP3s
  unsigned long long e[NQ + 1];
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    const uint32_t u = (uint32_t)occ[j] + bias;
    e[j] = 0ULL;
    if ((u & 0xFFFFFF88u) == 0u) e[j] = B.win_obs[u];   // one LOP3: column < 8 and row < 8; the table is zero outside 5 x 5
  }
  {
    const uint32_t u = o + bias;
    e[NQ] = 0ULL;
    if ((u & 0xFFFFFF88u) == 0u) e[NQ] = B.win_obs[u];
  }
  // PREDATOR (2) = PREY (3) with the low bit of the cell cleared
  return B.win_pos[p] | e[0] | e[1] | e[2] | (e[NQ] & 0xAAAAAAAAAAAAAAAAULL);
}
__device__ __forceinline__ void occ_of(uint32_t w0, uint32_t w1, int (&occ)[NQ]) {
#pragma unroll
  for (int j = 0; j < NQ; ++j) occ[j] = ((w0 >> (16 + j)) & 1u) ? SENT : (int)((w1 >> (8 * j)) & 255u);
}

__device__ __forceinline__ void initial_from_picks(const Blob& B, const int (&picks)[NP + NQ], Ents& e) {
  uint32_t used = 0;
  for (int i = 0; i < NP; ++i) {
    int k = 0, sel = 0;
    for (int c = 0; c < B.n_pred_start; ++c) {
      if ((used >> c) & 1u) continue;
      if (k == picks[i]) sel = c;
      ++k;
    }
    used |= 1u << sel;
    e.pred[i] = B.pred_start[sel];
  }
  used = 0;
  for (int i = 0; i < NQ; ++i) {
    int k = 0, sel = 0;
    for (int c = 0; c < B.n_prey_start; ++c) {
      if ((used >> c) & 1u) continue;
      if (k == picks[NP + i]) sel = c;
      ++k;
    }
    used |= 1u << sel;
    e.prey[i] = B.prey_start[sel];
  }
  e.caught = 0;
}

}  // namespace pp

// Adapter for the generic batch / search / belief kernels.  LUT = true: the table-driven transition /
// observation (Blob::lut_ok; the host picks the instantiation), same results.
template <bool LUT>
struct PPSimT {
  using Blob = pp::Blob;
  using Ctx = NoCtx;
  static constexpr int A = 5, OW = 1, SW = 2, NAG = 2, NW = 3, NPICK = 5;
  static constexpr bool ONE_WORD_NOISE = false;   // two Philox words per env-step (NOISE_PAIR below)
  struct State { uint32_t w[SW]; };
  __device__ static __forceinline__ void actions_from_word(uint32_t, int (&)[NAG]) {}

  // flattened-observation fields: the 25 window cells, two bits each (see DrivingSim::obs_field)
  static constexpr int NFIELD = 25;
  __host__ __device__ static constexpr uint32_t obs_field(int f) {
    return f < 16 ? (0u | ((uint32_t)(2 * f) << 8) | (2u << 16)) : (1u | ((uint32_t)(2 * (f - 16)) << 8) | (2u << 16));
  }

  // explicit-noise joint step (model.step): obs / rewards for every agent
  __device__ static __forceinline__ void step_all(const Blob& B, State& s, const int (&act)[NAG],
                                                  const uint32_t (&noise)[NW],
                                                  unsigned long long (&obs)[NAG][OW], double (&rew)[NAG],
                                                  uint32_t& flags, bool& was_live) {
    if constexpr (LUT) {
      was_live = ((s.w[0] >> 16) & 7u) != 7u;
      int occ[pp::NQ];
      const bool all_done = pp::transition_lut(B, s.w[0], s.w[1], act, noise, rew, occ);
      obs[0][0] = pp::observe_lut(B, s.w[0], occ, 0);
      obs[1][0] = pp::observe_lut(B, s.w[0], occ, 1);
      flags = all_done ? (0x80u | 3u) : 0u;
    } else {
      pp::Ents e;
      pp::unpack(s.w[0], s.w[1], e);
      was_live = e.caught != 7u;
      const bool all_done = pp::transition(B, e, act, noise, rew);
      pp::pack(e, s.w[0], s.w[1]);
      obs[0][0] = pp::observe(B, e, 0);
      obs[1][0] = pp::observe(B, e, 1);
      flags = all_done ? (0x80u | 3u) : 0u;
    }
  }
  // the same without observations (rollouts); term = per-agent terminated bits
  __device__ static __forceinline__ void step_noobs(const Blob& B, State& s, const int (&act)[NAG],
                                                    const uint32_t (&noise)[NW], double (&rew)[NAG],
                                                    uint32_t& term, bool& all_done) {
    if constexpr (LUT) {
      int occ[pp::NQ];
      all_done = pp::transition_lut(B, s.w[0], s.w[1], act, noise, rew, occ);
    } else {
      pp::Ents e;
      pp::unpack(s.w[0], s.w[1], e);
      all_done = pp::transition(B, e, act, noise, rew);
      pp::pack(e, s.w[0], s.w[1]);
    }
    term = all_done ? 3u : 0u;
  }
  // Step noise of a particle: HALF a Philox block.  Particles g and g + 32 (g with bit 5 clear) share the block with
  // counter ((g >> 6) << 5 | (g & 31), step, STREAM_STEP): the first takes words (x, y), the second (z, w).  Of a
  // pair (a, b): prey 0 uses a, prey 1 a * 0x10001 (high half = low + high half of a), prey 2 b, the uniform
  // joint action b * 0x10001.  (A prey needs a tie bit and one pick of at most five options from its word.)  The
  // fast kernel steps particles 32 j + lane in one thread, so it computes ONE block for two of its envs.
  static constexpr bool NOISE_PAIR = true;
  __device__ static __forceinline__ uint64_t noise_counter(uint64_t particle) {
    return ((particle >> 6) << 5) | (particle & 31u);
  }
  __device__ static __forceinline__ void noise_from_pair(uint32_t a, uint32_t b, int (&act)[NAG], uint32_t (&noise)[NW]) {
    noise[0] = a; noise[1] = a * 0x10001u; noise[2] = b;
    const uint32_t idx = uniform_below(b * 0x10001u, 25u);
    act[1] = (int)(idx / 5u);
    act[0] = (int)(idx - 5u * (uint32_t)act[1]);
  }
  template <class Seed>   // uint64_t seed or the PhiloxKeys of one
  __device__ static __forceinline__ void noise_from_philox(const Seed& seed, uint64_t particle, uint32_t step,
                                                           int (&act)[NAG], uint32_t (&noise)[NW]) {
    const uint4 w = philox_block(seed, noise_counter(particle), step, STREAM_STEP);
    const bool hi = (particle >> 5) & 1u;
    noise_from_pair(hi ? w.z : w.x, hi ? w.w : w.y, act, noise);
  }
  __device__ static __forceinline__ void noise_from_word(uint32_t word, uint32_t (&noise)[NW]) {
    const uint4 w = philox4x32_10(word, 0x9E3779B9u, 0x7F4A7C15u, 0x85EBCA6Bu, 0xC2B2AE35u, 0x27D4EB2Fu);
    noise[0] = w.x; noise[1] = w.y; noise[2] = w.z;
  }
  __device__ static __forceinline__ void observe_all(const Blob& B, const State& s, unsigned long long (&obs)[NAG][OW]) {
    if constexpr (LUT) {
      int occ[pp::NQ];
      pp::occ_of(s.w[0], s.w[1], occ);
      obs[0][0] = pp::observe_lut(B, s.w[0], occ, 0);
      obs[1][0] = pp::observe_lut(B, s.w[0], occ, 1);
    } else {
      pp::Ents e;
      pp::unpack(s.w[0], s.w[1], e);
      obs[0][0] = pp::observe(B, e, 0);
      obs[1][0] = pp::observe(B, e, 1);
    }
  }
  __device__ static __forceinline__ void initial_from_picks(const Blob& B, const int (&picks)[NPICK], State& s) {
    pp::Ents e;
    pp::initial_from_picks(B, picks, e);
    pp::pack(e, s.w[0], s.w[1]);
  }
  // picks from Philox: block (particle, 0, STREAM_INIT) words x,y -> predators; block
  // (particle, 0, STREAM_INIT2) words x,y,z -> prey
  __device__ static __forceinline__ void initial_state(const Blob& B, uint64_t seed, uint64_t particle, State& s) {
    const uint4 a = philox_block(seed, particle, 0u, STREAM_INIT);
    const uint4 b = philox_block(seed, particle, 0u, STREAM_INIT2);
    int picks[NPICK];
    picks[0] = (int)uniform_below(a.x, (uint32_t)B.n_pred_start);
    picks[1] = (int)uniform_below(a.y, (uint32_t)B.n_pred_start - 1u);
    picks[2] = (int)uniform_below(b.x, (uint32_t)B.n_prey_start);
    picks[3] = (int)uniform_below(b.y, (uint32_t)B.n_prey_start - 1u);
    picks[4] = (int)uniform_below(b.z, (uint32_t)B.n_prey_start - 2u);
    initial_from_picks(B, picks, s);
  }

  // ---- search / belief adapter ---------------------------------------------------
  __device__ static __forceinline__ void step(const Blob& B, const Ctx&, State& s, int ego, int ego_action,
                                              uint32_t word, uint32_t /*word2*/, unsigned long long (&obs)[OW],
                                              double& r, bool& done) {
    // `word` seeds the other predator's action and the three prey words
    const uint4 w = philox4x32_10(word, 0x9E3779B9u, 0x7F4A7C15u, 0x85EBCA6Bu, 0xC2B2AE35u, 0x27D4EB2Fu);
    int act[NAG];
    const int other = (int)uniform_below(w.w, 5u);
    act[0] = ego == 0 ? ego_action : other;
    act[1] = ego == 1 ? ego_action : other;
    const uint32_t noise[NW] = {w.x, w.y, w.z};
    double rew[NAG];
    if constexpr (LUT) {
      int occ[pp::NQ];
      done = pp::transition_lut(B, s.w[0], s.w[1], act, noise, rew, occ);
      obs[0] = pp::observe_lut(B, s.w[0], occ, ego);
    } else {
      pp::Ents e;
      pp::unpack(s.w[0], s.w[1], e);
      done = pp::transition(B, e, act, noise, rew);
      pp::pack(e, s.w[0], s.w[1]);
      obs[0] = pp::observe(B, e, ego);
    }
    r = rew[ego];
  }
  __device__ static __forceinline__ void rollout_step(const Blob& B, const Ctx&, State& s, int ego, uint32_t word,
                                                      double& r, bool& done) {
    const uint4 w = philox4x32_10(word, 0x9E3779B9u, 0x7F4A7C15u, 0x85EBCA6Bu, 0xC2B2AE35u, 0x27D4EB2Fu);
    int act[NAG];
    const uint32_t idx = uniform_below(w.w, 25u);
    act[1] = (int)(idx / 5u);
    act[0] = (int)(idx - 5u * (uint32_t)act[1]);
    const uint32_t noise[NW] = {w.x, w.y, w.z};
    double rew[NAG];
    if constexpr (LUT) {
      int occ[pp::NQ];
      done = pp::transition_lut(B, s.w[0], s.w[1], act, noise, rew, occ);
    } else {
      pp::Ents e;
      pp::unpack(s.w[0], s.w[1], e);
      done = pp::transition(B, e, act, noise, rew);
      pp::pack(e, s.w[0], s.w[1]);
    }
    r = rew[ego];
  }
  __device__ static __forceinline__ void observe(const Blob& B, const State& s, int ego, unsigned long long (&obs)[OW]) {
    if constexpr (LUT) {
      int occ[pp::NQ];
      pp::occ_of(s.w[0], s.w[1], occ);
      obs[0] = pp::observe_lut(B, s.w[0], occ, ego);
    } else {
      pp::Ents e;
      pp::unpack(s.w[0], s.w[1], e);
      obs[0] = pp::observe(B, e, ego);
    }
  }
};
using PPSim = PPSimT<false>;
using PPSimLut = PPSimT<true>;

}  // namespace posg
