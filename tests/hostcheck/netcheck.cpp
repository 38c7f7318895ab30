// Test infrastructure: the veneer's host-callable INetwork::evaluate (networks/INetwork.hpp:25-27) -- GridNetwork over a traced
// module on the GPU, RandomNetwork -- on positions reached by random legal play through the host GameNode.  Writes the
// positions and the answers as .npy files for tests/test_host_gpu.py, which recomputes them from the same module in PyTorch.
//   usage: netcheck <traced_model.pt> <number of positions> <seed> <output prefix>
#include "games/OthelloNode.hpp"
#include "networks/GridNetwork.hpp"
#include "networks/RandomNetwork.hpp"

#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

using namespace SPRL;
using State = OthelloNode::State;
constexpr int A = OTH_ACTION_SIZE;

int main(int argc, char** argv) {
    if (argc < 5) { std::fprintf(stderr, "usage: netcheck <model.pt> <positions> <seed> <output prefix>\n"); return 2; }
    const int want = std::atoi(argv[2]);
    std::mt19937_64 gen(std::strtoull(argv[3], nullptr, 10));
    const std::string prefix = argv[4];
    std::vector<State> states;
    std::vector<GameActionDist<A>> masks;
    std::vector<float> cells, players, maskRows;
    std::vector<std::unique_ptr<OthelloNode>> roots;
    while ((int)states.size() < want) {
        roots.push_back(std::make_unique<OthelloNode>());
        GameNode<OthelloNode, State, A>* cur = roots.back().get();
        while (!cur->isTerminal() && (int)states.size() < want) {
            states.push_back(cur->getGameState());
            masks.push_back(cur->getActionMask());
            for (int i = 0; i < OTH_BOARD_SIZE; ++i) cells.push_back((float)cur->cells()[i]);
            players.push_back((float)(int)cur->getPlayer());
            std::vector<int> legal;
            for (int a = 0; a < A; ++a) { maskRows.push_back(cur->getActionMask()[a]); if (cur->getActionMask()[a] > 0.0f) legal.push_back(a); }
            cur = cur->getAddChild((ActionIdx)legal[gen() % legal.size()]);
        }
    }
    GridNetwork<OTH_BOARD_WIDTH, OTH_BOARD_WIDTH, OTH_HISTORY_SIZE, A> net(argv[1]);
    const auto got = net.evaluate(states, masks);
    RandomNetwork<State, A> uniform;
    const auto flat = uniform.evaluate(states, masks);
    std::vector<float> policy, value, upolicy;
    for (size_t b = 0; b < got.size(); ++b) {
        for (int a = 0; a < A; ++a) { policy.push_back(got[b].first[a]); upolicy.push_back(flat[b].first[a]); }
        value.push_back(got[b].second);
        if (flat[b].second != 0.0f) { std::fprintf(stderr, "RandomNetwork value is not 0\n"); return 1; }
    }
    const uint64_t n = states.size();
    writeNpy(prefix + "_cells.npy", cells, { n, (uint64_t)OTH_BOARD_SIZE });
    writeNpy(prefix + "_player.npy", players, { n });
    writeNpy(prefix + "_mask.npy", maskRows, { n, (uint64_t)A });
    writeNpy(prefix + "_policy.npy", policy, { n, (uint64_t)A });
    writeNpy(prefix + "_value.npy", value, { n });
    writeNpy(prefix + "_uniform.npy", upolicy, { n, (uint64_t)A });
    std::printf("netcheck ok: %llu positions, %d + %d evaluations counted\n", (unsigned long long)n, net.getNumEvals(), uniform.getNumEvals());
    return 0;
}
