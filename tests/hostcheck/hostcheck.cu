// hostcheck.cu -- TEST ONLY.  Runs the HOST instantiation of the product's
// __host__ __device__ bitboard rules (sprl_b200/csrc/games.cuh, rng.cuh) on the
// CPU so that the rules can be compared with the oracle where no GPU exists.
// It is built by tests/test_host_rules.py, never shipped, never loaded by the
// product; the GPU path is tested separately (-m gpu) through the C ABI.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../sprl_b200/csrc/games.cuh"
#include "../../sprl_b200/csrc/rng.cuh"

using namespace sprl;

template <int W>
struct VecHist {
    const std::vector<Bits<W>>* b0;
    const std::vector<Bits<W>>* b1;
    bool seen(const Bits<W>& x0, const Bits<W>& x1) const {
        for (size_t i = 0; i < b0->size(); ++i) if ((*b0)[i] == x0 && (*b1)[i] == x1) return true;
        return false;
    }
    template <class F> void for_each(F&& f) const {
        for (size_t i = 0; i < b0->size(); ++i) f((*b0)[i], (*b1)[i]);
    }
};

template <class G>
static uint64_t perft_rec(const typename G::P& p, int depth, std::vector<Bits<G::W>>& h0, std::vector<Bits<G::W>>& h1) {
    if (depth == 0 || p.terminal) return 1;
    uint64_t total = 0;
    int n = p.n_legal();
    for (int k = 0; k < n; ++k) {
        typename G::P nx;
        VecHist<G::W> hist = { &h0, &h1 };
        G::next(p, legal_action<G>(p, k), hist, nx);
        h0.push_back(p.b[0]); h1.push_back(p.b[1]);
        total += perft_rec<G>(nx, depth - 1, h0, h1);
        h0.pop_back(); h1.pop_back();
    }
    return total;
}

template <class G>
static int64_t rollout(uint64_t seed, uint64_t first_game, int ngames, int64_t cap, int32_t* game_steps,
                       int8_t* cells, int8_t* player, int8_t* terminal, int8_t* winner, int8_t* mask, int32_t* action) {
    int64_t pos = 0;
    for (int g = 0; g < ngames; ++g) {
        Rng rng = { seed, first_game + (uint64_t)g, 0 };
        typename G::P cur, nx;
        G::start(cur);
        std::vector<Bits<G::W>> h0, h1;
        int steps = 0;
        for (;;) {
            if (pos >= cap) return -1;
            for (int i = 0; i < G::CELLS; ++i) cells[pos * G::CELLS + i] = cur.b[0].test(i) ? 0 : (cur.b[1].test(i) ? 1 : -1);
            player[pos] = cur.player; terminal[pos] = cur.terminal;
            winner[pos] = cur.winner == WINNER_ZERO ? 0 : (cur.winner == WINNER_ONE ? 1 : -1);
            for (int a = 0; a < G::ACTIONS; ++a) {
                bool ok = (G::HAS_PASS && a == G::CELLS) ? cur.pass_legal : cur.legal.test(a);
                mask[pos * G::ACTIONS + a] = ok;
            }
            ++steps;
            if (cur.terminal) { action[pos++] = -1; break; }
            int pick = legal_action<G>(cur, rng_uniform_int(rng, 0, cur.n_legal() - 1));
            action[pos++] = pick;
            VecHist<G::W> hist = { &h0, &h1 };
            G::next(cur, pick, hist, nx);
            h0.push_back(cur.b[0]); h1.push_back(cur.b[1]);
            cur = nx;
        }
        game_steps[g] = steps;
    }
    return pos;
}

#define DISPATCH(game, CALL)                              \
    switch (game) {                                       \
    case 0: { typedef Othello G; return CALL; }           \
    case 1: { typedef ConnectFour G; return CALL; }       \
    case 2: { typedef Go<7> G; return CALL; }             \
    case 3: { typedef Go<9> G; return CALL; }             \
    default: return -1;                                   \
    }

template <class G> static int64_t perft_entry(int depth, uint64_t* count) {
    typename G::P s;
    G::start(s);
    std::vector<Bits<G::W>> h0, h1;
    *count = perft_rec<G>(s, depth, h0, h1);
    return 0;
}

extern "C" {
int64_t hostcheck_perft(int game, int depth, uint64_t* count) { DISPATCH(game, perft_entry<G>(depth, count)); }
int64_t hostcheck_rollout(int game, uint64_t seed, uint64_t first_game, int ngames, int64_t cap, int32_t* game_steps,
                          int8_t* cells, int8_t* player, int8_t* terminal, int8_t* winner, int8_t* mask, int32_t* action) {
    DISPATCH(game, rollout<G>(seed, first_game, ngames, cap, game_steps, cells, player, terminal, winner, mask, action));
}
uint32_t hostcheck_philox(uint64_t seed, uint64_t game, uint64_t ctr) { return philox_word0(seed, game, ctr); }
float hostcheck_powf(float x, float e) { return det_powf(x, e); }
float hostcheck_expf(float x) { return det_expf(x); }
float hostcheck_gamma(uint64_t seed, uint64_t game, uint64_t* ctr, float alpha) {
    Rng r = { seed, game, *ctr };
    float g = rng_gamma(r, alpha);
    *ctr = r.ctr;
    return g;
}
}
