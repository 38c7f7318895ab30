// Test infrastructure: the veneer's host GameNode / GridState (games/GameNode.hpp:50-200, games/GridState.hpp:56-115 of the
// reference) against the C ABI they are built on.  For every game: random lines of legal moves are played node by node
// through getAddChild and, independently, as ONE sprl_env_line call; cells, player, mask, terminal flag and winner must agree
// at every position, children are cached, parents linked, the GridState carries the last HISTORY boards (newest first), an
// illegal action throws, pruning drops the siblings.  Then selfPlay() (selfplay/SelfPlay.hpp:50-192, one game on the device
// returned as GridStates / action distributions / outcomes) for the three games: the samples of a move are its state under
// every symmetry in index order, which the HOST functions ISymmetrizer::symmetrizeState / symmetrizeActionDist
// (symmetry/ISymmetrizer.hpp:33-56) must reproduce exactly from the identity sample; histories chain from move to move.
// Prints "gamenode ok" on success.   usage: gamenode [seed]
#include "games/ConnectFourNode.hpp"
#include "games/GoNode.hpp"
#include "games/OthelloNode.hpp"
#include "networks/RandomNetwork.hpp"
#include "selfplay/SelfPlay.hpp"
#include "symmetry/ConnectFourSymmetrizer.hpp"
#include "symmetry/D4GridSymmetrizer.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

using namespace SPRL;

#define EXPECT(cond)                                                                         \
    do {                                                                                     \
        if (!(cond)) { std::fprintf(stderr, "gamenode: %s failed (line %d)\n", #cond, __LINE__); std::exit(1); } \
    } while (0)

template <typename Node, int ACTION_SIZE>
static int walk(std::mt19937_64& gen, int games) {
    using State = typename Node::State;
    int positions = 0;
    for (int g = 0; g < games; ++g) {
        Node root;
        GameNode<Node, State, ACTION_SIZE>* cur = &root;
        std::vector<int32_t> line;
        std::vector<const GameNode<Node, State, ACTION_SIZE>*> path { cur };
        while (!cur->isTerminal()) {
            std::vector<int> legal;
            for (int a = 0; a < ACTION_SIZE; ++a) if (cur->getActionMask()[a] > 0.0f) legal.push_back(a);
            EXPECT(!legal.empty());
            const int a = legal[gen() % legal.size()];
            // an action outside the mask is refused
            for (int b = 0; b < ACTION_SIZE; ++b)
                if (!(cur->getActionMask()[b] > 0.0f)) {
                    bool thrown = false;
                    try { cur->getAddChild((ActionIdx)b); } catch (const EngineError& e) { thrown = e.code == SPRL_E_INVALID; }
                    EXPECT(thrown);
                    break;
                }
            Node* next = cur->getAddChild((ActionIdx)a);
            EXPECT(next == cur->getAddChild((ActionIdx)a));            // cached
            EXPECT(next->getParent() == cur);
            EXPECT(cur->getWinner() == Player::NONE);
            line.push_back(a);
            cur = next;
            path.push_back(cur);
        }
        // the same line in one call
        const size_t n = line.size() + 1;
        std::vector<int8_t> cells(n * State::BOARD), player(n), terminal(n), winner(n), mask(n * ACTION_SIZE);
        check(sprl_env_line(currentDevice(), Node::GAME, (int32_t)line.size(), line.data(), cells.data(), player.data(),
                            terminal.data(), winner.data(), mask.data()));
        for (size_t k = 0; k < n; ++k) {
            const auto* node = path[k];
            for (int i = 0; i < State::BOARD; ++i) EXPECT(node->cells()[i] == cells[k * State::BOARD + i]);
            EXPECT((int)node->getPlayer() == player[k]);
            EXPECT(node->isTerminal() == (terminal[k] != 0));
            EXPECT((int)node->getWinner() == winner[k]);
            for (int a = 0; a < ACTION_SIZE; ++a) EXPECT((node->getActionMask()[a] > 0.0f) == (mask[k * ACTION_SIZE + a] != 0));
            // GridState: the last HISTORY boards, newest first
            const State st = node->getGameState();
            EXPECT(st.getPlayer() == node->getPlayer());
            EXPECT(st.size() == (int)std::min<size_t>(k + 1, State::HISTORY));
            for (int t = 0; t < st.size(); ++t)
                for (int i = 0; i < State::BOARD; ++i) EXPECT((int8_t)st.getHistory()[t][i] == cells[(k - t) * State::BOARD + i]);
            ++positions;
        }
        const auto rewards = cur->getRewards();
        EXPECT(rewards[0] == -rewards[1]);
        EXPECT((cur->getWinner() == Player::ZERO) == (rewards[0] == 1.0f));
        // pruning keeps one child
        if (line.size() >= 1) {
            root.pruneChildrenExcept((ActionIdx)line[0]);
            EXPECT(root.getAddChild((ActionIdx)line[0]) == path[1]);
        }
    }
    return positions;
}

template <typename State>
static bool sameState(const State& a, const State& b) {
    if (a.size() != b.size() || a.getPlayer() != b.getPlayer()) return false;
    for (int t = 0; t < a.size(); ++t)
        for (int i = 0; i < State::BOARD; ++i) if (a.getHistory()[t][i] != b.getHistory()[t][i]) return false;
    return true;
}

template <typename Node, int ACTION_SIZE, typename Symmetrizer>
static int selfPlayCheck(Symmetrizer& sym, int sims) {
    using State = typename Node::State;
    RandomNetwork<State, ACTION_SIZE> net;
    auto [states, dists, outcomes] = selfPlay<Node, State, ACTION_SIZE>(std::make_unique<Node>(), &net, sims, 4, 2, 0.25f, 0.3f,
                                                                      InitQ::PARENT, &sym, true);
    const int S = sym.numSymmetries();
    EXPECT(!states.empty() && states.size() % S == 0 && states.size() == dists.size() && states.size() == outcomes.size());
    std::vector<SymmetryIdx> all;
    for (int k = 0; k < S; ++k) all.push_back((SymmetryIdx)k);
    const size_t moves = states.size() / S;
    Node start;
    for (size_t m = 0; m < moves; ++m) {
        const State& base = states[m * S];                                // symmetry 0 is the identity
        EXPECT(base.size() == (int)std::min<size_t>(m + 1, State::HISTORY));
        if (m == 0) {
            EXPECT(base.getPlayer() == Player::ZERO);
            for (int i = 0; i < State::BOARD; ++i) EXPECT((int8_t)base.getHistory()[0][i] == start.cells()[i]);
        } else {
            const State& prev = states[(m - 1) * S];                      // the history chains: board t of move m = board t-1 of move m-1
            for (int t = 1; t < base.size(); ++t)
                for (int i = 0; i < State::BOARD; ++i) EXPECT(base.getHistory()[t][i] == prev.getHistory()[t - 1][i]);
        }
        const std::vector<State> hs = sym.symmetrizeState(base, all);
        const auto hd = sym.symmetrizeActionDist(dists[m * S], all);
        float sum = 0.0f;
        for (int a = 0; a < ACTION_SIZE; ++a) sum += dists[m * S][a];
        EXPECT(std::fabs(sum - 1.0f) < 1e-5f);
        for (int k = 0; k < S; ++k) {
            EXPECT(sameState(hs[k], states[m * S + k]));
            for (int a = 0; a < ACTION_SIZE; ++a) EXPECT(hd[k][a] == dists[m * S + k][a]);
            EXPECT(outcomes[m * S + k] == outcomes[m * S]);
            EXPECT(outcomes[m * S] == 1.0f || outcomes[m * S] == -1.0f || outcomes[m * S] == 0.0f);
            // a symmetry followed by its inverse is the identity
            const std::vector<SymmetryIdx> back { sym.inverseSymmetry((SymmetryIdx)k) };
            EXPECT(sameState(sym.symmetrizeState(hs[k], back)[0], base));
        }
    }
    return (int)moves;
}

int main(int argc, char** argv) {
    std::mt19937_64 gen(argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 1);
    const int oth = walk<OthelloNode, OTH_ACTION_SIZE>(gen, 3);
    const int c4 = walk<ConnectFourNode, C4_ACTION_SIZE>(gen, 6);
    const int go = walk<GoNode, GO_ACTION_SIZE>(gen, 2);
    D4GridSymmetrizer<OTH_BOARD_WIDTH, OTH_HISTORY_SIZE> othSym;
    ConnectFourSymmetrizer c4Sym;
    D4GridSymmetrizer<GO_BOARD_WIDTH, GO_HISTORY_SIZE> goSym;
    const int m0 = selfPlayCheck<OthelloNode, OTH_ACTION_SIZE>(othSym, 32);
    const int m1 = selfPlayCheck<ConnectFourNode, C4_ACTION_SIZE>(c4Sym, 32);
    const int m2 = selfPlayCheck<GoNode, GO_ACTION_SIZE>(goSym, 24);
    std::printf("selfPlay ok: %d Othello, %d Connect Four, %d Go moves, every symmetric sample reproduced on the host\n", m0, m1, m2);
    std::printf("gamenode ok: %d Othello, %d Connect Four, %d Go positions\n", oth, c4, go);
    return 0;
}
