// ref_trace.cpp -- TEST INFRASTRUCTURE.  Driver around the UNMODIFIED reference
// sources (compiled where they lie under /root/reference/cpp/src, see
// oracle/Makefile) that emits golden traces for the hot path:
//
//   perft    <game> <maxdepth>
//   rollout  <game> <seed> <first_game> <ngames> <out.trace>
//   selfplay <game> <evaluator> <seed> <first_game> <ngames> <sims> <batch>
//            <queue> <eps> <alpha> <noise 0|1> <sym 0|1> <parent|zero|drop> <out.trace>
//   match    <game> <evaluator0> <evaluator1> <seed> <first_game> <ngames> <sims> <batch> <queue>
//            <sym0 0|1> <initq0> <sym1 0|1> <initq1> <out.trace>
//
//   treewalk <game> <evaluator> <seed> <first_game> <ngames> <sims> <batch> <queue> <eps> <alpha> <noise 0|1> <sym 0|1>
//            <parent|zero|drop> <out.trace>    ONE tree per game driven through the public UCTTree API by a caller
//                                              that plays the first most-visited action: searchAndGetLeaves /
//                                              evaluateAndBackpropLeaves until <sims>, getDecisionNode()'s statistics,
//                                              advanceDecision(action) -- the trace of the step-wise C ABI
//   npy      x <out.npy> <d0> [<d1> ...]       the array value[i] = 0.25 * i - 3 of that shape through the reference's own
//                                              npy::write_npy (utils/npy.hpp:616-639), as selfplay/GridWorker.hpp:173-196 calls it
//
// evaluator = hash | hash1 (a second, independent HashNet) | uniform (networks/RandomNetwork.hpp)
//           | heuristic (networks/OthelloHeuristic.cpp, Othello only)
//           | pt:<file> (builds with -DSPRL_REF_WITH_TORCH only, `make ref_torch`: the reference's LibTorch
//             GridNetwork, networks/GridNetwork.hpp:37-145, on a TorchScript file -- its own exp / mask / sum / divide).
//
// game = othello | c4 | go.  Randomness comes from random_shim.cpp (the
// contract stream of oracle/oracle_rng.h).  `selfplay` runs every game twice:
// once through the reference's own SPRL::selfPlay (selfplay/SelfPlay.hpp:50-192)
// and once through an instrumented loop over the reference's public UCTTree API
// (uct/UCTTree.hpp:38-210) that records the root statistics after every search;
// it aborts unless both produce identical samples and consume the same number
// of draws, so the recorded statistics are those of the reference's selfPlay.
// `match` does the same for the match-play path of Evaluate.cpp:93-157: the
// reference's own UCTNetworkAgent (agents/UCTNetworkAgent.hpp:42-108) + playGame
// (evaluate/play.hpp:24-69), then an instrumented loop that must reproduce its
// actions, winner and draw count.
#include "games/ConnectFourNode.hpp"
#include "games/GoNode.hpp"
#include "games/OthelloNode.hpp"
#include "agents/UCTNetworkAgent.hpp"
#include "evaluate/play.hpp"
#include "networks/OthelloHeuristic.hpp"
#include "networks/RandomNetwork.hpp"
#include "selfplay/SelfPlay.hpp"
#include "symmetry/ConnectFourSymmetrizer.hpp"
#include "symmetry/D4GridSymmetrizer.hpp"

#include "utils/npy.hpp"
#ifdef SPRL_REF_WITH_TORCH
#include "networks/GridNetwork.hpp"
#endif

#include "hashnet.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <type_traits>

extern "C" void sprl_shim_set_stream(uint64_t seed, uint64_t game);
extern "C" uint64_t sprl_shim_get_counter();

using namespace SPRL;

// ---- trace file: a sequence of named raw arrays ---------------------------
// record = u32 name_len | name | u8 dtype ('b' i8,'i' i32,'f' f32,'Q' u64) |
//          u32 ndim | u64 dims[ndim] | raw little-endian data
struct TraceWriter {
    std::ofstream out;
    explicit TraceWriter(const std::string& path) : out(path, std::ios::binary) {
        if (!out) { std::cerr << "cannot open " << path << "\n"; std::exit(2); }
    }
    template <typename T>
    void put(const std::string& name, char dtype, const std::vector<T>& data, std::vector<uint64_t> dims) {
        uint64_t n = 1;
        for (uint64_t d : dims) n *= d;
        if (n != data.size()) { std::cerr << "shape mismatch for " << name << "\n"; std::exit(2); }
        uint32_t nl = (uint32_t)name.size(), nd = (uint32_t)dims.size();
        out.write((const char*)&nl, 4);
        out.write(name.data(), nl);
        out.write(&dtype, 1);
        out.write((const char*)&nd, 4);
        out.write((const char*)dims.data(), 8 * nd);
        out.write((const char*)data.data(), sizeof(T) * data.size());
    }
};

// ---- per-game compile-time description -----------------------------------
struct OthelloDesc {
    using Node = OthelloNode;
    static constexpr int R = OTH_BOARD_WIDTH, C = OTH_BOARD_WIDTH, H = OTH_HISTORY_SIZE, A = OTH_ACTION_SIZE;
    using Sym = D4GridSymmetrizer<OTH_BOARD_WIDTH, OTH_HISTORY_SIZE>;
};
struct C4Desc {
    using Node = ConnectFourNode;
    static constexpr int R = C4_NUM_ROWS, C = C4_NUM_COLS, H = C4_HISTORY_SIZE, A = C4_ACTION_SIZE;
    using Sym = ConnectFourSymmetrizer;
};
struct GoDesc {
    using Node = GoNode;
    static constexpr int R = GO_BOARD_WIDTH, C = GO_BOARD_WIDTH, H = GO_HISTORY_SIZE, A = GO_ACTION_SIZE;
    using Sym = D4GridSymmetrizer<GO_BOARD_WIDTH, GO_HISTORY_SIZE>;
};

template <class D> using StateOf = GridState<D::R * D::C, D::H>;
template <class D> using GNode = GameNode<typename D::Node, StateOf<D>, D::A>;

// ---- perft -----------------------------------------------------------------
template <class D>
uint64_t perft(GNode<D>* node, int depth) {
    if (depth == 0 || node->isTerminal()) return 1;
    uint64_t total = 0;
    for (int a = 0; a < D::A; ++a) {
        if (node->getActionMask()[a] == 0.0f) continue;
        GNode<D>* child = node->getAddChild((ActionIdx)a);
        total += perft<D>(child, depth - 1);
        node->pruneChildrenExcept((ActionIdx)a);  // frees the previous sibling's subtree
    }
    return total;
}

template <class D>
int cmdPerft(int maxDepth) {
    for (int d = 1; d <= maxDepth; ++d) {
        typename D::Node root;
        std::cout << d << " " << perft<D>(&root, d) << std::endl;
    }
    return 0;
}

// ---- random rollouts -------------------------------------------------------
template <class D>
int cmdRollout(uint64_t seed, uint64_t firstGame, int nGames, const std::string& outPath) {
    constexpr int B = D::R * D::C;
    std::vector<int32_t> gameSteps;       // positions recorded per game (incl. start and terminal)
    std::vector<int8_t> cells, player, terminal, winner, mask;
    std::vector<int32_t> action;
    for (int g = 0; g < nGames; ++g) {
        sprl_shim_set_stream(seed, firstGame + g);
        typename D::Node root;
        GNode<D>* cur = &root;
        int steps = 0;
        for (;;) {
            StateOf<D> st = cur->getGameState();
            for (int i = 0; i < B; ++i) cells.push_back((int8_t)st.getHistory()[0][i]);
            player.push_back((int8_t)cur->getPlayer());
            terminal.push_back(cur->isTerminal() ? 1 : 0);
            winner.push_back((int8_t)cur->getWinner());
            std::vector<int> legal;
            for (int a = 0; a < D::A; ++a) {
                bool ok = cur->getActionMask()[a] != 0.0f;
                mask.push_back(ok ? 1 : 0);
                if (ok) legal.push_back(a);
            }
            ++steps;
            if (cur->isTerminal()) { action.push_back(-1); break; }
            int pick = legal[GetRandom().UniformInt(0, (int)legal.size() - 1)];
            action.push_back(pick);
            cur = cur->getAddChild((ActionIdx)pick);
        }
        gameSteps.push_back(steps);
    }
    uint64_t P = player.size();
    TraceWriter w(outPath);
    w.put("game_steps", 'i', gameSteps, { (uint64_t)nGames });
    w.put("cells", 'b', cells, { P, (uint64_t)B });
    w.put("player", 'b', player, { P });
    w.put("terminal", 'b', terminal, { P });
    w.put("winner", 'b', winner, { P });
    w.put("mask", 'b', mask, { P, (uint64_t)D::A });
    w.put("action", 'i', action, { P });
    std::cout << "{\"positions\": " << P << "}" << std::endl;
    return 0;
}

// ---- self-play ---------------------------------------------------------------
// Embedding of sample states into float planes: restated from
// selfplay/GridWorker.hpp:146-171 (own stones, opponent stones per history
// step; zero padding for missing history; colour plane).
template <class D>
void embedStates(const std::vector<StateOf<D>>& states, std::vector<float>& out) {
    constexpr int B = D::R * D::C;
    for (const auto& s : states) {
        Piece own = pieceFromPlayer(s.getPlayer());
        Piece opp = otherPiece(own);
        for (int t = 0; t < D::H; ++t) {
            for (Piece which : { own, opp }) {
                for (int i = 0; i < B; ++i) {
                    out.push_back((t < s.size() && s.getHistory()[t][i] == which) ? 1.0f : 0.0f);
                }
            }
        }
        for (int i = 0; i < B; ++i) out.push_back(s.getPlayer() == Player::ZERO ? 1.0f : 0.0f);
    }
}

// ---- evaluators and options by name -----------------------------------------
template <class D>
std::unique_ptr<INetwork<StateOf<D>, D::A>> makeEvaluator(const std::string& kind) {
    using State = StateOf<D>;
    if (kind == "hash") return std::make_unique<SPRLREF::HashNet<D::R * D::C, D::H, D::A>>(0);
    if (kind == "hash1") return std::make_unique<SPRLREF::HashNet<D::R * D::C, D::H, D::A>>(1);
    if (kind == "uniform") return std::make_unique<RandomNetwork<State, D::A>>();
#ifdef SPRL_REF_WITH_TORCH
    if (kind.rfind("pt:", 0) == 0) {
        torch::set_num_threads(1);
        return std::make_unique<GridNetwork<D::R, D::C, D::H, D::A>>(kind.substr(3));
    }
#endif
    if constexpr (std::is_same_v<typename D::Node, OthelloNode>) {
        if (kind == "heuristic") return std::make_unique<OthelloHeuristic>();
    }
    std::cerr << "unknown evaluator " << kind << "\n";
    std::exit(2);
}

static InitQ parseInitQ(const std::string& s) {
    if (s == "zero") return InitQ::ZERO;
    if (s == "parent") return InitQ::PARENT;
    if (s == "drop") return InitQ::DROP_PARENT;
    std::cerr << "unknown init-Q " << s << "\n";
    std::exit(2);
}

template <class D>
int cmdSelfplay(const std::string& evalKind, uint64_t seed, uint64_t firstGame, int nGames,
                int sims, int maxBatch, int maxQueue, float eps, float alpha,
                bool addNoise, bool useSym, InitQ initQ, const std::string& outPath) {
    constexpr int B = D::R * D::C;
    constexpr int A = D::A;
    using State = StateOf<D>;
    using Dist = GameActionDist<A>;

    auto netOwner = makeEvaluator<D>(evalKind);
    INetwork<State, A>* net = netOwner.get();

    typename D::Sym symObj;
    ISymmetrizer<State, A>* sym = useSym ? &symObj : nullptr;

    std::vector<int32_t> gameMoves, gameSamples;
    std::vector<uint64_t> gameCtr;
    std::vector<float> mvN, mvW, mvP, mvRootN, mvRootW;
    std::vector<int32_t> mvAction, mvTrav, mvEvals;
    std::vector<int8_t> mvPlayer, mvBoard;
    std::vector<float> smpStates, smpDists, smpOutcomes;

    std::streambuf* coutBuf = std::cout.rdbuf();

    for (int g = 0; g < nGames; ++g) {
        // Pass 1: the reference's own selfPlay.
        sprl_shim_set_stream(seed, firstGame + g);
        auto [refStates, refDists, refOutcomes] = selfPlay<typename D::Node, State, A>(
            std::make_unique<typename D::Node>(), net, sims, maxBatch, maxQueue,
            eps, alpha, initQ, sym, addNoise);
        uint64_t refCtr = sprl_shim_get_counter();

        // Pass 2: instrumented loop over the public UCTTree API.
        sprl_shim_set_stream(seed, firstGame + g);
        std::vector<SymmetryIdx> allSym;
        if (sym) for (int i = 0; i < sym->numSymmetries(); ++i) allSym.push_back((SymmetryIdx)i);
        UCTTree<typename D::Node, State, A> tree { std::make_unique<typename D::Node>(),
                                                   eps, alpha, initQ, sym, addNoise };
        std::vector<State> states;
        std::vector<Dist> dists;
        std::vector<Player> players;
        int moveCount = 0;
        while (!tree.getDecisionNode()->isTerminal()) {
            auto* root = const_cast<UCTNode<typename D::Node, State, A>*>(tree.getDecisionNode());
            State rootState = root->getGameState();
            if (sym) {
                auto ss = sym->symmetrizeState(rootState, allSym);
                states.insert(states.end(), ss.begin(), ss.end());
            } else {
                states.push_back(rootState);
            }
            int evals0 = net->getNumEvals();
            int trav = 0;
            while (trav < sims) {
                auto [leaves, t] = tree.searchAndGetLeaves(maxBatch, maxQueue, net, U_WEIGHT);
                if (!leaves.empty()) tree.evaluateAndBackpropLeaves(leaves, net);
                trav += t;
            }
            const auto* es = root->getEdgeStatistics();
            for (int a = 0; a < A; ++a) {
                mvN.push_back(es->m_numVisits[a]);
                mvW.push_back(es->m_totalValues[a]);
                mvP.push_back(es->m_childPriors[a]);
            }
            mvRootN.push_back(root->N());
            mvRootW.push_back(root->W());
            mvTrav.push_back(trav);
            mvEvals.push_back(net->getNumEvals() - evals0);
            mvPlayer.push_back((int8_t)root->getPlayer());
            for (int i = 0; i < B; ++i) mvBoard.push_back((int8_t)rootState.getHistory()[0][i]);

            Dist visits = es->m_numVisits;
            Dist pdf = visits / visits.sum();
            pdf = (moveCount < EARLY_GAME_CUTOFF) ? pdf.pow(EARLY_GAME_EXP) : pdf.pow(REST_GAME_EXP);
            pdf = pdf / pdf.sum();
            Dist cdf = pdf.cumsum();
            cdf = cdf / cdf[A - 1];
            if (sym) {
                auto dd = sym->symmetrizeActionDist(pdf, allSym);
                dists.insert(dists.end(), dd.begin(), dd.end());
            } else {
                dists.push_back(pdf);
            }
            int action = GetRandom().SampleCDF(std::vector<float>(cdf.begin(), cdf.end()));
            mvAction.push_back(action);
            players.push_back(root->getPlayer());
            tree.advanceDecision((ActionIdx)action);
            ++moveCount;
        }
        std::array<Value, 2> rewards = tree.getDecisionNode()->getRewards();
        std::vector<float> outcomes;
        int perMove = sym ? sym->numSymmetries() : 1;
        for (Player p : players)
            for (int i = 0; i < perMove; ++i) outcomes.push_back(rewards[(int)p]);
        uint64_t myCtr = sprl_shim_get_counter();

        // The instrumented loop must be indistinguishable from selfPlay.
        bool same = (myCtr == refCtr) && states.size() == refStates.size() &&
                    dists.size() == refDists.size() && outcomes == refOutcomes;
        for (size_t i = 0; same && i < states.size(); ++i) {
            same = states[i].size() == refStates[i].size() && states[i].getPlayer() == refStates[i].getPlayer();
            for (int t = 0; same && t < states[i].size(); ++t)
                same = states[i].getHistory()[t] == refStates[i].getHistory()[t];
            for (int a = 0; same && a < A; ++a)
                same = std::memcmp(&dists[i].begin()[a], &refDists[i].begin()[a], 4) == 0;
        }
        if (!same) {
            std::cout.rdbuf(coutBuf);
            std::cerr << "instrumented loop diverged from reference selfPlay in game " << g << "\n";
            return 3;
        }

        gameMoves.push_back(moveCount);
        gameSamples.push_back((int32_t)states.size());
        gameCtr.push_back(myCtr);
        embedStates<D>(states, smpStates);
        for (const auto& d : dists) for (int a = 0; a < A; ++a) smpDists.push_back(d[a]);
        smpOutcomes.insert(smpOutcomes.end(), outcomes.begin(), outcomes.end());
    }

    uint64_t M = mvAction.size(), S = smpOutcomes.size();
    TraceWriter w(outPath);
    w.put("game_moves", 'i', gameMoves, { (uint64_t)nGames });
    w.put("game_samples", 'i', gameSamples, { (uint64_t)nGames });
    w.put("game_rng_draws", 'Q', gameCtr, { (uint64_t)nGames });
    w.put("move_N", 'f', mvN, { M, (uint64_t)A });
    w.put("move_W", 'f', mvW, { M, (uint64_t)A });
    w.put("move_P", 'f', mvP, { M, (uint64_t)A });
    w.put("move_root_N", 'f', mvRootN, { M });
    w.put("move_root_W", 'f', mvRootW, { M });
    w.put("move_action", 'i', mvAction, { M });
    w.put("move_traversals", 'i', mvTrav, { M });
    w.put("move_evals", 'i', mvEvals, { M });
    w.put("move_player", 'b', mvPlayer, { M });
    w.put("move_board", 'b', mvBoard, { M, (uint64_t)B });
    w.put("states", 'f', smpStates, { S, (uint64_t)(2 * D::H + 1), (uint64_t)D::R, (uint64_t)D::C });
    w.put("distributions", 'f', smpDists, { S, (uint64_t)A });
    w.put("outcomes", 'f', smpOutcomes, { S });
    std::cout << "{\"moves\": " << M << ", \"samples\": " << S << "}" << std::endl;
    return 0;
}

// ---- match play (Evaluate.cpp:93-157) ----------------------------------------
template <class D>
int cmdMatch(const std::string& evalKind0, const std::string& evalKind1, uint64_t seed, uint64_t firstGame, int nGames,
             int sims, int maxBatch, int maxQueue, bool sym0, InitQ initQ0, bool sym1, InitQ initQ1,
             const std::string& outPath) {
    constexpr int B = D::R * D::C;
    constexpr int A = D::A;
    using State = StateOf<D>;
    using Tree = UCTTree<typename D::Node, State, A>;
    using Agent = UCTNetworkAgent<typename D::Node, State, A>;

    auto net0 = makeEvaluator<D>(evalKind0);
    auto net1 = makeEvaluator<D>(evalKind1);
    typename D::Sym symObj;
    const float eps = 0.25f, alpha = 0.1f;                 // Evaluate.cpp:97-98,106-107

    std::vector<int32_t> gameMoves, gameWinner, gameFirst;
    std::vector<uint64_t> gameCtr;
    std::vector<float> mvN, mvW, mvP, mvRootN, mvRootW;
    std::vector<int32_t> mvAction, mvTrav, mvAgent;
    std::vector<int8_t> mvPlayer, mvBoard;

    for (int g = 0; g < nGames; ++g) {
        const uint64_t t = firstGame + g;                  // the loop index of Evaluate.cpp:93
        // Pass 1: the reference's agents and playGame.
        sprl_shim_set_stream(seed, t);
        std::vector<int> refActions;
        Player refWinner;
        {
            Tree tree0 { std::make_unique<typename D::Node>(), eps, alpha, initQ0, sym0 ? &symObj : nullptr, true };
            Tree tree1 { std::make_unique<typename D::Node>(), eps, alpha, initQ1, sym1 ? &symObj : nullptr, true };
            Agent agent0 { net0.get(), &tree0, sims, maxBatch, maxQueue };
            Agent agent1 { net1.get(), &tree1, sims, maxBatch, maxQueue };
            // forwards to the reference's agent and notes what it played
            struct Recorder : IAgent<typename D::Node, State, A> {
                const IAgent<typename D::Node, State, A>* inner;
                std::vector<int>* log;
                ActionIdx act(const GameNode<typename D::Node, State, A>* node, bool verbose = false) const override {
                    ActionIdx a = inner->act(node, verbose);
                    log->push_back((int)a);
                    return a;
                }
                void opponentAct(const ActionIdx action) const override { inner->opponentAct(action); }
            };
            Recorder rec0, rec1;
            rec0.inner = &agent0; rec0.log = &refActions;
            rec1.inner = &agent1; rec1.log = &refActions;
            std::array<IAgent<typename D::Node, State, A>*, 2> agents;
            if (t % 2 == 0) agents = { &rec0, &rec1 }; else agents = { &rec1, &rec0 };
            typename D::Node rootNode {};
            refWinner = playGame<typename D::Node, State, A>(&rootNode, agents, false);
        }
        uint64_t refCtr = sprl_shim_get_counter();

        // Pass 2: instrumented loop over the public UCTTree API.
        sprl_shim_set_stream(seed, t);
        Tree tree0 { std::make_unique<typename D::Node>(), eps, alpha, initQ0, sym0 ? &symObj : nullptr, true };
        Tree tree1 { std::make_unique<typename D::Node>(), eps, alpha, initQ1, sym1 ? &symObj : nullptr, true };
        Tree* trees[2] = { &tree0, &tree1 };
        INetwork<State, A>* nets[2] = { net0.get(), net1.get() };
        std::vector<int> actions;
        typename D::Node gameRoot {};
        GNode<D>* cur = &gameRoot;
        while (!tree0.getDecisionNode()->isTerminal()) {
            int player = (int)tree0.getDecisionNode()->getPlayer();
            int agent = (t % 2 == 0) ? player : 1 - player;        // which network moves
            Tree& tree = *trees[agent];
            auto* root = const_cast<UCTNode<typename D::Node, State, A>*>(tree.getDecisionNode());
            State rootState = root->getGameState();
            int trav = 0;
            while (trav < sims) {
                auto [leaves, n] = tree.searchAndGetLeaves(maxBatch, maxQueue, nets[agent]);   // default uWeight, as the agent
                if (!leaves.empty()) tree.evaluateAndBackpropLeaves(leaves, nets[agent]);
                trav += n;
            }
            const auto* es = root->getEdgeStatistics();
            for (int a = 0; a < A; ++a) {
                mvN.push_back(es->m_numVisits[a]);
                mvW.push_back(es->m_totalValues[a]);
                mvP.push_back(es->m_childPriors[a]);
            }
            mvRootN.push_back(root->N());
            mvRootW.push_back(root->W());
            mvTrav.push_back(trav);
            mvAgent.push_back(agent);
            mvPlayer.push_back((int8_t)player);
            for (int i = 0; i < B; ++i) mvBoard.push_back((int8_t)rootState.getHistory()[0][i]);
            auto visits = es->m_numVisits;
            int action = (int)std::distance(visits.begin(), std::max_element(visits.begin(), visits.end()));
            mvAction.push_back(action);
            actions.push_back(action);
            tree.advanceDecision((ActionIdx)action);
            trees[1 - agent]->advanceDecision((ActionIdx)action);
            cur = cur->getAddChild((ActionIdx)action);
        }
        Player winner = cur->getWinner();
        uint64_t myCtr = sprl_shim_get_counter();
        if (myCtr != refCtr || winner != refWinner || actions != refActions) {
            std::cerr << "instrumented match loop diverged from reference playGame in game " << g << "\n";
            return 3;
        }
        gameMoves.push_back((int32_t)actions.size());
        gameWinner.push_back((int32_t)winner);
        gameFirst.push_back((int32_t)(t % 2));               // network that plays Player::ZERO
        gameCtr.push_back(myCtr);
    }
    uint64_t M = mvAction.size();
    TraceWriter w(outPath);
    w.put("game_moves", 'i', gameMoves, { (uint64_t)nGames });
    w.put("game_winner", 'i', gameWinner, { (uint64_t)nGames });
    w.put("game_first", 'i', gameFirst, { (uint64_t)nGames });
    w.put("game_rng_draws", 'Q', gameCtr, { (uint64_t)nGames });
    w.put("move_N", 'f', mvN, { M, (uint64_t)A });
    w.put("move_W", 'f', mvW, { M, (uint64_t)A });
    w.put("move_P", 'f', mvP, { M, (uint64_t)A });
    w.put("move_root_N", 'f', mvRootN, { M });
    w.put("move_root_W", 'f', mvRootW, { M });
    w.put("move_action", 'i', mvAction, { M });
    w.put("move_traversals", 'i', mvTrav, { M });
    w.put("move_agent", 'i', mvAgent, { M });
    w.put("move_player", 'b', mvPlayer, { M });
    w.put("move_board", 'b', mvBoard, { M, (uint64_t)B });
    std::cout << "{\"moves\": " << M << "}" << std::endl;
    return 0;
}

// ---- treewalk: the public UCTTree API with the caller choosing the moves --------------------------
template <class D>
int cmdTreewalk(const std::string& evalKind, uint64_t seed, uint64_t firstGame, int nGames, int sims, int maxBatch, int maxQueue,
                float eps, float alpha, bool addNoise, bool useSym, InitQ initQ, const std::string& outPath) {
    constexpr int A = D::A;
    using State = StateOf<D>;
    auto net = makeEvaluator<D>(evalKind);
    typename D::Sym symObj;
    std::vector<int32_t> gameMoves, gameWinner, mvAction, mvTrav;
    std::vector<uint64_t> gameCtr;
    std::vector<float> mvN, mvW, mvP, mvRootN, mvRootW;
    std::vector<int8_t> mvPlayer;
    for (int g = 0; g < nGames; ++g) {
        sprl_shim_set_stream(seed, firstGame + g);
        UCTTree<typename D::Node, State, A> tree { std::make_unique<typename D::Node>(), eps, alpha, initQ, useSym ? &symObj : nullptr, addNoise };
        int moves = 0;
        typename D::Node gameRoot {};
        GNode<D>* cur = &gameRoot;              // the same game on a plain game tree, for the winner
        while (!tree.getDecisionNode()->isTerminal()) {
            auto* root = const_cast<UCTNode<typename D::Node, State, A>*>(tree.getDecisionNode());
            int trav = 0;
            while (trav < sims) {
                auto [leaves, n] = tree.searchAndGetLeaves(maxBatch, maxQueue, net.get(), U_WEIGHT);
                if (!leaves.empty()) tree.evaluateAndBackpropLeaves(leaves, net.get());
                trav += n;
            }
            const auto* es = root->getEdgeStatistics();
            for (int a = 0; a < A; ++a) {
                mvN.push_back(es->m_numVisits[a]);
                mvW.push_back(es->m_totalValues[a]);
                mvP.push_back(es->m_childPriors[a]);
            }
            mvRootN.push_back(root->N());
            mvRootW.push_back(root->W());
            mvTrav.push_back(trav);
            mvPlayer.push_back((int8_t)root->getPlayer());
            auto visits = es->m_numVisits;
            int action = (int)std::distance(visits.begin(), std::max_element(visits.begin(), visits.end()));
            mvAction.push_back(action);
            tree.advanceDecision((ActionIdx)action);
            cur = cur->getAddChild((ActionIdx)action);
            ++moves;
        }
        gameMoves.push_back(moves);
        gameWinner.push_back((int32_t)cur->getWinner());
        gameCtr.push_back(sprl_shim_get_counter());
    }
    uint64_t M = mvAction.size();
    TraceWriter w(outPath);
    w.put("game_moves", 'i', gameMoves, { (uint64_t)nGames });
    w.put("game_winner", 'i', gameWinner, { (uint64_t)nGames });
    w.put("game_rng_draws", 'Q', gameCtr, { (uint64_t)nGames });
    w.put("move_N", 'f', mvN, { M, (uint64_t)A });
    w.put("move_W", 'f', mvW, { M, (uint64_t)A });
    w.put("move_P", 'f', mvP, { M, (uint64_t)A });
    w.put("move_root_N", 'f', mvRootN, { M });
    w.put("move_root_W", 'f', mvRootW, { M });
    w.put("move_action", 'i', mvAction, { M });
    w.put("move_traversals", 'i', mvTrav, { M });
    w.put("move_player", 'b', mvPlayer, { M });
    std::cout << "{\"moves\": " << M << "}" << std::endl;
    return 0;
}

template <class D>
int dispatch(int argc, char** argv) {
    std::string cmd = argv[1];
    if (cmd == "perft" && argc == 4) return cmdPerft<D>(std::atoi(argv[3]));
    if (cmd == "rollout" && argc == 7)
        return cmdRollout<D>(std::strtoull(argv[3], 0, 10), std::strtoull(argv[4], 0, 10), std::atoi(argv[5]), argv[6]);
    if (cmd == "selfplay" && argc == 16) {
        InitQ q = parseInitQ(argv[14]);
        return cmdSelfplay<D>(argv[3], std::strtoull(argv[4], 0, 10), std::strtoull(argv[5], 0, 10),
                              std::atoi(argv[6]), std::atoi(argv[7]), std::atoi(argv[8]), std::atoi(argv[9]),
                              (float)std::atof(argv[10]), (float)std::atof(argv[11]),
                              std::atoi(argv[12]) != 0, std::atoi(argv[13]) != 0, q, argv[15]);
    }
    if (cmd == "treewalk" && argc == 16) {
        InitQ q = parseInitQ(argv[14]);
        return cmdTreewalk<D>(argv[3], std::strtoull(argv[4], 0, 10), std::strtoull(argv[5], 0, 10),
                              std::atoi(argv[6]), std::atoi(argv[7]), std::atoi(argv[8]), std::atoi(argv[9]),
                              (float)std::atof(argv[10]), (float)std::atof(argv[11]),
                              std::atoi(argv[12]) != 0, std::atoi(argv[13]) != 0, q, argv[15]);
    }
    if (cmd == "match" && argc == 16)
        return cmdMatch<D>(argv[3], argv[4], std::strtoull(argv[5], 0, 10), std::strtoull(argv[6], 0, 10), std::atoi(argv[7]),
                           std::atoi(argv[8]), std::atoi(argv[9]), std::atoi(argv[10]), std::atoi(argv[11]) != 0,
                           parseInitQ(argv[12]), std::atoi(argv[13]) != 0, parseInitQ(argv[14]), argv[15]);
    std::cerr << "bad arguments; see the header of ref_trace.cpp\n";
    return 2;
}

// ---- npy: the reference's writer on a known array ------------------------------------------------
static int cmdNpy(int argc, char** argv) {
    npy::npy_data_ptr<float> d {};
    size_t n = 1;
    for (int i = 4; i < argc; ++i) { d.shape.push_back(std::strtoul(argv[i], 0, 10)); n *= d.shape.back(); }
    std::vector<float> v(n);
    for (size_t i = 0; i < n; ++i) v[i] = 0.25f * (float)i - 3.0f;
    d.data_ptr = v.data();
    npy::write_npy(argv[3], d);
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 5 && std::string(argv[1]) == "npy") return cmdNpy(argc, argv);
    if (argc < 3) { std::cerr << "usage: ref_trace <perft|rollout|selfplay> <othello|c4|go> ...\n"; return 2; }
    std::string game = argv[2];
    if (game == "othello") return dispatch<OthelloDesc>(argc, argv);
    if (game == "c4") return dispatch<C4Desc>(argc, argv);
    if (game == "go") return dispatch<GoDesc>(argc, argv);
    std::cerr << "unknown game " << game << "\n";
    return 2;
}
