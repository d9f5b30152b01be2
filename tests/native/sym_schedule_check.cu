// Host-side check of the work decomposition of the sweeps (csrc/nbody_sym.cuh, csrc/nbody_kernels.cuh), compiled and run by
// tests/test_host.py::test_symmetric_schedule_visits_every_block_pair_once — no GPU needed.
//
// For many (N, IBLK, TILE, shard count) it replays what the kernels do with the geometry helpers
// (sym_ncols, sym_tiles_in_block, sym_row_tiles, sym_advance) and checks the properties the physics needs:
//   1. every unordered pair of body-blocks {A, B} (A == B included) is visited exactly once over all shards,
//   2. sym_row_tiles (the host's row_start table) equals the number of tiles the walker steps through,
//   3. the walker enumerates (row, column, tile) in exactly the order of the nested loops and ends at the
//      end of the last local row, whatever flat offset a CTA starts from (binary search + walk, as in the kernel),
//   4. the j-tiles of a row cover its column blocks completely (ragged last block included),
//   5. block rows carry equal work up to one column block (load balance across shards),
//   6. CTA ranges cut at chunk granularity (sym_cta_range, and cost-weighted: sym_locate_weighted) tile the share's
//      (tile, chunk) items exactly once; the weighted ranges carry equal cost and start with the right walker.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../gravitation_b200/csrc/nbody_sym.cuh"

using namespace gravb200;

static long long fails = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (fails < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } ++fails; } } while (0)

template <int IBLK, int TILE>
void check(long long n, int world, long long chunk) {
    const int Bt = (int)((n + IBLK - 1) / IBLK);
    std::vector<int> visited((size_t)Bt * Bt, 0);
    long long min_row = -1, max_row = -1;
    for (int rank = 0; rank < world; ++rank) {
        const long long row0 = std::min<long long>((long long)rank * chunk, n);
        const long long n_local = std::max<long long>(0, std::min<long long>(chunk, n - row0));
        if (n_local == 0) continue;
        const int nib = (int)((n_local + IBLK - 1) / IBLK);
        SymParams p;
        p.n_total = n; p.row0 = row0; p.n_local = n_local; p.n_iblocks = nib; p.n_gblocks = Bt; p.gblock0 = (int)(row0 / IBLK);
        std::vector<long long> rs((size_t)nib + 1, 0);
        for (int i = 0; i < nib; ++i) rs[i + 1] = rs[i] + sym_row_tiles(n, IBLK, TILE, Bt, p.gblock0 + i);
        // nested-loop enumeration, the reference order
        std::vector<SymWalker> order;
        for (int I = 0; I < nib; ++I) {
            const int Ig = p.gblock0 + I;
            const int nc = sym_ncols(Bt, Ig);
            long long tiles = 0, bodies = 0;
            for (int c = 0; c < nc; ++c) {
                const int K = (Ig + c) % Bt;
                ++visited[(size_t)std::min(Ig, K) * Bt + std::max(Ig, K)];
                CHECK(c == 0 || K != Ig, "row %d revisits its diagonal block", Ig);
                const int tb = sym_tiles_in_block(n, IBLK, TILE, K);
                for (int t = 0; t < tb; ++t) {
                    order.push_back({I, c, t});
                    const long long j0 = (long long)K * IBLK + (long long)t * TILE;
                    bodies += std::min<long long>(TILE, n - j0);
                    CHECK(j0 < n, "tile starts past the last body");
                }
                tiles += tb;
                const long long in_block = std::min<long long>(IBLK, n - (long long)K * IBLK);
                CHECK(in_block > 0, "empty column block");
            }
            CHECK(tiles == rs[I + 1] - rs[I], "sym_row_tiles %lld != walked %lld (n=%lld Ig=%d)", rs[I + 1] - rs[I], tiles, n, Ig);
            long long expect = 0;
            for (int c = 0; c < nc; ++c) expect += std::min<long long>(IBLK, n - (long long)((Ig + c) % Bt) * IBLK);
            CHECK(bodies == expect, "tiles of row %d cover %lld bodies, blocks hold %lld", Ig, bodies, expect);
            if (min_row < 0 || nc < min_row) min_row = nc;
            if (nc > max_row) max_row = nc;
        }
        CHECK((long long)order.size() == rs[nib], "flat size");
        // the walker from the start, and from every 7th flat offset located the kernel's way
        for (long long lo = 0; lo < rs[nib]; lo += (lo == 0 ? 1 : 7)) {
            int a = 0, b = nib;
            while (b - a > 1) { const int m = (a + b) >> 1; if (rs[m] <= lo) a = m; else b = m; }
            SymWalker w{a, 0, 0};
            long long rem = lo - rs[a];
            for (;;) {
                const int tb = sym_tiles_in_block(n, IBLK, TILE, (p.gblock0 + a + w.c) % Bt);
                if (rem < tb) break;
                rem -= tb; ++w.c;
            }
            w.t = (int)rem;
            const long long span = lo == 0 ? rs[nib] : std::min<long long>(rs[nib] - lo, 40);
            for (long long k = 0; k < span; ++k) {
                const SymWalker& e = order[(size_t)(lo + k)];
                CHECK(w.I == e.I && w.c == e.c && w.t == e.t, "walker (%d,%d,%d) != (%d,%d,%d) at %lld (n=%lld)", w.I, w.c, w.t, e.I, e.c, e.t, lo + k, n);
                sym_advance<IBLK, TILE>(w, p);
            }
            if (lo == 0) CHECK(w.I == nib && w.c == 0 && w.t == 0, "walker ends at (%d,%d,%d), not at row %d", w.I, w.c, w.t, nib);
        }
    }
    for (int A = 0; A < Bt; ++A)
        for (int B = A; B < Bt; ++B)
            CHECK(visited[(size_t)A * Bt + B] == 1, "block pair (%d,%d) of %d visited %d times (n=%lld iblk=%d world=%d)", A, B, Bt, visited[(size_t)A * Bt + B], n, IBLK, world);
    CHECK(max_row - min_row <= 1, "block rows differ by %lld column blocks", max_row - min_row);
}

template <int IBLK, int TILE>
void sweep(long long& cases) {
    const long long sizes[] = {1, 2, TILE - 1, TILE, TILE + 1, IBLK - 1, IBLK, IBLK + 1, 2LL * IBLK, 2LL * IBLK + 5, 3LL * IBLK - 1, 5LL * IBLK + TILE,
                               8LL * IBLK, 9LL * IBLK - 7, 17LL * IBLK + 33, 40000, 65536, 100003, 262144, 1048576};
    for (long long n : sizes) {
        check<IBLK, TILE>(n, 1, n); ++cases;
        for (int world : {2, 3, 4, 8}) {
            const long long plain = (n + world - 1) / world;
            const long long chunk = (plain + IBLK - 1) / IBLK * IBLK;   // shards are whole blocks (gravb200_partition)
            check<IBLK, TILE>(n, world, chunk); ++cases;
        }
    }
}

// stream-K partition of the ordered sweep (csrc/nbody_kernels.cuh): CTA c owns flat tiles [sk_lo(c), sk_lo(c+1));
// the ranges tile [0, total) without gap or overlap, differ by at most one tile, and sk_owner inverts sk_lo
void check_stream_k(long long& cases) {
    const long long totals[] = {1, 2, 3, 7, 147, 148, 149, 295, 296, 297, 1000, 4096, 65536, 1048576, 4194304 + 17, (1LL << 31) + 5, 3LL << 33};
    const long long grids[] = {1, 2, 3, 37, 148, 296, 592};
    for (long long total : totals)
        for (long long S : grids) {
            ++cases;
            CHECK(sk_lo(total, 0, S) == 0 && sk_lo(total, S, S) == total, "ends (total=%lld S=%lld)", total, S);
            long long lo_min = total, lo_max = 0;
            for (long long c = 0; c < S; ++c) {
                const long long lo = sk_lo(total, c, S), hi = sk_lo(total, c + 1, S);
                CHECK(lo <= hi, "range order");
                lo_min = std::min(lo_min, hi - lo); lo_max = std::max(lo_max, hi - lo);
                if (hi > lo) {
                    CHECK(sk_owner(total, lo, S) == c, "owner of first tile %lld of CTA %lld is %lld (total=%lld S=%lld)", lo, c, sk_owner(total, lo, S), total, S);
                    CHECK(sk_owner(total, hi - 1, S) == c, "owner of last tile %lld of CTA %lld is %lld (total=%lld S=%lld)", hi - 1, c, sk_owner(total, hi - 1, S), total, S);
                }
            }
            CHECK(lo_max - lo_min <= 1, "ranges differ by %lld tiles (total=%lld S=%lld)", lo_max - lo_min, total, S);
        }
}

// speed-proportional shares of the flat list (sym_share_bounds): for random published (items, ns) the P shares tile
// [0, total) exactly — every shard computes its own bounds, hi of shard r must BE lo of shard r + 1 — follow the
// speeds (items / ns) to one item, and fall back to the equal shares while any shard has not published yet
static void check_shares(long long& cases) {
    unsigned long long seed = 12345;
    auto rnd = [&]() { seed = seed * 6364136223846793005ull + 1442695040888963407ull; return seed >> 33; };
    for (int rep = 0; rep < 4000; ++rep) {
        const int P = 1 + (int)(rnd() % 16);
        const long long total = 1 + (long long)(rnd() % 3000000);
        unsigned long long stats[2 * (kMaxPeers + 1)] = {};
        double speed[kMaxPeers + 1], sum = 0;
        const bool missing = rep % 7 == 0;
        for (int q = 0; q < P; ++q) {
            stats[q] = 1 + rnd() % 100000;
            stats[(kMaxPeers + 1) + q] = 1000000 + rnd() % 40000000;
            speed[q] = (double)stats[q] / (double)stats[(kMaxPeers + 1) + q];
            sum += speed[q];
        }
        if (missing) stats[rnd() % P] = 0;
        long long prev_hi = 0;
        for (int me = 0; me < P; ++me) {
            long long lo, hi;
            const long long eq_lo = sk_lo(total, me, P), eq_hi = sk_lo(total, me + 1, P);
            sym_share_bounds(total, P, me, stats, eq_lo, eq_hi, lo, hi);
            CHECK(lo == prev_hi, "share %d of %d starts at %lld, the previous one ended at %lld", me, P, lo, prev_hi);
            CHECK(hi >= lo && hi <= total, "share %d: [%lld, %lld) of %lld", me, lo, hi, total);
            if (missing) CHECK(lo == eq_lo && hi == eq_hi, "a shard has not published: equal shares expected");
            else CHECK(std::abs((double)(hi - lo) - (double)total * speed[me] / sum) <= 2.0, "share %d is %lld items, its speed asks for %.1f", me, hi - lo, (double)total * speed[me] / sum);
            prev_hi = hi;
        }
        CHECK(prev_hi == total, "the shares end at %lld of %lld", prev_hi, total);
        ++cases;
    }
}

// CTA ranges at chunk granularity (sym_cta_range<CHUNKS, 1>), replayed the way the kernels use them: tiles [lo, hi)
// of the share, chunks [c_first, CHUNKS) of the first tile, [0, c_last) of the last one ([c_first, c_last) when they
// are the same tile).  Every (tile, chunk) of the share belongs to exactly one CTA, no range is empty, the CTAs'
// chunk counts differ by at most one, and SPLIT = 0 reproduces the whole-tile stream-K ranges.
template <int CHUNKS>
static void check_split(long long& cases) {
    const long long totals[] = {1, 2, 3, 5, 9, 37, 110, 144, 147, 148, 149, 264, 295, 1518, 4224, 22446, 351918};
    const long long grids[] = {1, 2, 3, 37, 110, 148, 296};
    const long long share_los[] = {0, 7, 123456};
    for (long long total : totals)
        for (long long S : grids)
            for (long long share_lo : share_los) {
                ++cases;
                std::vector<int> seen((size_t)total * CHUNKS, 0);
                long long cmin = total * CHUNKS, cmax = 0, busy = 0;
                for (long long b = 0; b < S; ++b) {
                    SymRange rg, r0;
                    const bool any0 = sym_cta_range<CHUNKS, 0>(share_lo, total, b, S, r0);
                    CHECK(any0 == (sk_lo(total, b, S) < sk_lo(total, b + 1, S)), "whole-tile range: emptiness");
                    if (any0) CHECK(r0.lo == share_lo + sk_lo(total, b, S) && r0.hi == share_lo + sk_lo(total, b + 1, S) && r0.c_first == 0 && r0.c_last == CHUNKS,
                                    "whole-tile range of CTA %lld (total=%lld S=%lld)", b, total, S);
                    if (!sym_cta_range<CHUNKS, 1>(share_lo, total, b, S, rg)) continue;
                    ++busy;
                    CHECK(rg.lo >= share_lo && rg.hi <= share_lo + total && rg.lo < rg.hi, "tiles [%lld, %lld) outside the share", rg.lo, rg.hi);
                    CHECK(rg.c_first >= 0 && rg.c_first < CHUNKS && rg.c_last >= 1 && rg.c_last <= CHUNKS, "chunk bounds %d, %d", rg.c_first, rg.c_last);
                    const int ntiles = (int)(rg.hi - rg.lo);
                    long long mine = 0;
                    for (int k = 0; k < ntiles; ++k) {
                        const int cb = k == 0 ? rg.c_first : 0, ce = k == ntiles - 1 ? rg.c_last : CHUNKS;
                        CHECK(cb < ce, "empty chunk range [%d, %d) in tile %d of %d (total=%lld S=%lld b=%lld)", cb, ce, k, ntiles, total, S, b);
                        for (int c = cb; c < ce; ++c) { ++seen[(size_t)(rg.lo - share_lo + k) * CHUNKS + c]; ++mine; }
                    }
                    cmin = std::min(cmin, mine); cmax = std::max(cmax, mine);
                }
                for (size_t i = 0; i < seen.size(); ++i)
                    if (seen[i] != 1) { CHECK(false, "chunk %zu of tile %zu visited %d times (total=%lld S=%lld CHUNKS=%d)", i % CHUNKS, i / CHUNKS, seen[i], total, S, CHUNKS); break; }
                CHECK(busy == std::min<long long>(S, total * CHUNKS), "%lld busy CTAs of %lld for %lld chunks", busy, S, total * CHUNKS);
                CHECK(cmax - cmin <= 1, "CTA ranges differ by %lld chunks (total=%lld S=%lld)", cmax - cmin, total, S);
            }
}

// Cost-weighted chunk-granular ranges (sym_locate_weighted): over the shards' equal shares of the tile list and
// many grid sizes, the CTAs' chunk ranges tile every share exactly once, the walker a CTA starts with is the
// nested-loop position of its first tile, and the CTAs' COSTS (diagonal chunks w_diag, symmetric chunks w_sym)
// differ by less than two symmetric chunks.
template <int IBLK, int TILE>
static void check_weighted(long long n, int world, int w_sym, int w_diag, long long& cases) {
    constexpr int CH = TILE / 32;
    const int Bt = (int)((n + IBLK - 1) / IBLK);
    std::vector<long long> rs((size_t)Bt + 1, 0), rc((size_t)Bt + 1, 0);
    std::vector<SymWalker> order;
    std::vector<int> is_diag;
    for (int I = 0; I < Bt; ++I) {
        const long long row_tiles = sym_row_tiles(n, IBLK, TILE, Bt, I);
        rs[I + 1] = rs[I] + row_tiles;
        rc[I + 1] = rc[I] + sym_cost_in_row(n, IBLK, TILE, I, row_tiles, w_sym, w_diag);
        for (int c = 0; c < sym_ncols(Bt, I); ++c)
            for (int t = 0; t < sym_tiles_in_block(n, IBLK, TILE, (I + c) % Bt); ++t) { order.push_back({I, c, t}); is_diag.push_back(c == 0); }
    }
    const long long total = rs[Bt];
    CHECK((long long)order.size() == total, "flat size");
    SymParams p;
    memset(&p, 0, sizeof(p));
    p.n_total = n; p.row0 = 0; p.n_local = n; p.n_iblocks = Bt; p.n_gblocks = Bt; p.gblock0 = 0;
    p.row_start = rs.data(); p.row_cost = rc.data(); p.w_sym = w_sym; p.w_diag = w_diag;
    auto cost_at = [&](long long t) -> long long {
        if (t >= total) return rc[Bt];
        int a = 0, b = Bt;
        while (b - a > 1) { const int m = (a + b) >> 1; if (rs[m] <= t) a = m; else b = m; }
        return rc[a] + sym_cost_in_row(n, IBLK, TILE, a, t - rs[a], w_sym, w_diag);
    };
    for (long long S : {1LL, 3LL, 37LL, 148LL, 296LL}) {
        ++cases;
        std::vector<int> seen((size_t)total * CH, 0);
        for (int rank = 0; rank < world; ++rank) {
            const long long lo = sk_lo(total, rank, world), hi = sk_lo(total, rank + 1, world);
            p.item_lo = lo; p.item_hi = hi; p.cost_lo = cost_at(lo); p.cost_hi = cost_at(hi);
            long long cmin = -1, cmax = -1;
            for (long long b = 0; b < S; ++b) {
                SymRange rg; SymWalker w0;
                if (!sym_locate_weighted<IBLK, TILE>(p, b, S, rg, w0)) continue;
                CHECK(rg.lo >= lo && rg.hi <= hi && rg.lo < rg.hi, "tiles [%lld, %lld) outside the share [%lld, %lld)", rg.lo, rg.hi, lo, hi);
                const SymWalker& e = order[(size_t)rg.lo];
                CHECK(w0.I == e.I && w0.c == e.c && w0.t == e.t, "walker (%d,%d,%d) != (%d,%d,%d) at tile %lld (n=%lld S=%lld)", w0.I, w0.c, w0.t, e.I, e.c, e.t, rg.lo, n, S);
                const int ntiles = (int)(rg.hi - rg.lo);
                long long cost = 0;
                for (int k = 0; k < ntiles; ++k) {
                    const int cb = k == 0 ? rg.c_first : 0, ce = k == ntiles - 1 ? rg.c_last : CH;
                    CHECK(cb < ce, "empty chunk range [%d, %d) (n=%lld S=%lld b=%lld)", cb, ce, n, S, b);
                    for (int c = cb; c < ce; ++c) { ++seen[(size_t)(rg.lo + k) * CH + c]; cost += is_diag[(size_t)(rg.lo + k)] ? w_diag : w_sym; }
                }
                if (cmin < 0 || cost < cmin) cmin = cost;
                if (cost > cmax) cmax = cost;
            }
            if ((hi - lo) * CH >= S) CHECK(cmax - cmin < 2 * std::max(w_sym, w_diag), "CTA costs differ by %lld (n=%lld iblk=%d S=%lld world=%d rank=%d)", cmax - cmin, n, IBLK, S, world, rank);
        }
        for (size_t i = 0; i < seen.size(); ++i)
            if (seen[i] != 1) { CHECK(false, "chunk %zu of tile %zu visited %d times (n=%lld iblk=%d tile=%d S=%lld world=%d)", i % CH, i / CH, seen[i], n, IBLK, TILE, S, world); break; }
    }
}

template <int IBLK, int TILE>
static void sweep_weighted(long long& cases) {
    const long long sizes[] = {1, 33, TILE + 1, IBLK - 1, IBLK, IBLK + 1, 2LL * IBLK + 5, 5000, 9600, 12001, 16384, 24576, 40000, 65536, 100003};
    for (long long n : sizes)
        for (int world : {1, 2, 8}) {
            check_weighted<IBLK, TILE>(n, world, 4, 3, cases);
            check_weighted<IBLK, TILE>(n, world, 5, 4, cases);
        }
}

int main() {
    long long cases = 0;
    sweep_weighted<3072, 512>(cases); sweep_weighted<2048, 512>(cases); sweep_weighted<2048, 256>(cases); sweep_weighted<1024, 256>(cases); sweep_weighted<1024, 128>(cases);
    check_stream_k(cases);
    check_shares(cases);
    check_split<16>(cases); check_split<8>(cases); check_split<4>(cases);
    sweep<3072, 512>(cases); sweep<3072, 256>(cases); sweep<2048, 512>(cases); sweep<2048, 256>(cases);
    sweep<1024, 256>(cases); sweep<1536, 256>(cases); sweep<1536, 128>(cases); sweep<1024, 128>(cases);
    printf("%s: %lld cases, %lld failed checks\n", fails ? "FAILED" : "OK", cases, fails);
    return fails ? 1 : 0;
}
