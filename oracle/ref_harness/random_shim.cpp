// random_shim.cpp -- TEST INFRASTRUCTURE.  Link-time replacement for the
// reference's cpp/src/utils/random.cpp when building oracle/_ref.  The
// reference header utils/random.hpp is used untouched; only the out-of-line
// members (random.cpp:29-98) are re-defined here so that every draw comes from
// the counter-based contract stream in oracle/oracle_rng.h.
//
// The reference's call order within one game is preserved because these are the
// very functions its own code calls (uct/UCTNode.hpp:250, :332;
// uct/UCTTree.hpp:146; selfplay/SelfPlay.hpp:140; utils/Zobrist.hpp:39).
#include "utils/random.hpp"

#include "../oracle_rng.h"

#include <vector>

namespace {
// Constant-initialised so that static-init users (GoNode::s_zobrist) are safe.
// Until a driver selects a game stream, draws come from a reserved stream.
orng_t g_stream = { 0x5a6f627269737431ULL, ~0ULL, 0 };
}  // namespace

extern "C" void sprl_shim_set_stream(uint64_t seed, uint64_t game) {
    g_stream.seed = seed;
    g_stream.game = game;
    g_stream.ctr = 0;
}
extern "C" uint64_t sprl_shim_get_counter() { return g_stream.ctr; }
extern "C" orng_t* sprl_shim_stream() { return &g_stream; }

namespace SPRL {

Random& GetRandom() {
    static Random random(1, 1);
    return random;
}

Random::Random(uint64_t seed, int stream) : seed_(seed), impl_(seed, stream) {}

void Random::Dirichlet(float alpha, std::vector<float>& samples) {
    orng_dirichlet(&g_stream, alpha, samples.data(), (int)samples.size());
}

int Random::UniformInt(int a, int b) { return orng_uniform_int(&g_stream, a, b); }

uint64_t Random::UniformUint64(uint64_t, uint64_t) { return orng_u64(&g_stream); }

int Random::SampleCDF(const std::vector<float>& cdf) {
    return orng_sample_cdf(&g_stream, cdf.data(), (int)cdf.size());
}

}  // namespace SPRL
