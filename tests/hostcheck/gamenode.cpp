// Test infrastructure: the veneer's host GameNode / GridState (games/GameNode.hpp:50-200, games/GridState.hpp:56-115 of the
// reference) against the C ABI they are built on.  For every game: random lines of legal moves are played node by node
// through getAddChild and, independently, as ONE sprl_env_line call; cells, player, mask, terminal flag and winner must agree
// at every position, children are cached, parents linked, the GridState carries the last HISTORY boards (newest first), an
// illegal action throws, pruning drops the siblings.  Prints "gamenode ok" on success.   usage: gamenode [seed]
#include "games/ConnectFourNode.hpp"
#include "games/GoNode.hpp"
#include "games/OthelloNode.hpp"

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

int main(int argc, char** argv) {
    std::mt19937_64 gen(argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 1);
    const int oth = walk<OthelloNode, OTH_ACTION_SIZE>(gen, 3);
    const int c4 = walk<ConnectFourNode, C4_ACTION_SIZE>(gen, 6);
    const int go = walk<GoNode, GO_ACTION_SIZE>(gen, 2);
    std::printf("gamenode ok: %d Othello, %d Connect Four, %d Go positions\n", oth, c4, go);
    return 0;
}
