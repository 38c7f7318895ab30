// veneer.hpp -- header-only C++ host layer over the C ABI (include/sprl_b200.h) that keeps
// the names and signatures of the reference's driver API, so that a worker main written
// against willwin4sure/sprl (cpp/src/OTHWorker.cpp, C4Worker.cpp, GoWorker.cpp) compiles
// against this tree with only the include path changed:
//
//   SPRL::Player / Piece / ActionIdx / Value          games/GameNode.hpp:17-38, games/GridState.hpp:18-55
//   SPRL::GridState<B,H>, GameActionDist<A>           tags: states live on the device as bitboards
//   SPRL::OthelloNode / ConnectFourNode / GoNode      tags naming the device rules (games.cuh)
//   SPRL::INetwork<State,A>, RandomNetwork            networks/INetwork.hpp, RandomNetwork.hpp
//   SPRL::ISymmetrizer<State,A>, D4GridSymmetrizer,   symmetry/*.hpp
//         ConnectFourSymmetrizer
//   SPRL::InitQ                                       uct/UCTNode.hpp:24-28
//   SPRL::runIteration<Impl,State,A>(...), selfPlay   selfplay/SelfPlay.hpp:203-208, :50-192
//   SPRL::waitModelPath, SPRL::runWorker<NN,Impl,R,C,H,A>(...)   selfplay/GridWorker.hpp:35-55,84-91
//   SPRL::playMatch<Impl,State,A>(...)                the game loop of Evaluate.cpp:93-157 (UCTNetworkAgent::act /
//                                                     opponentAct inside playGame), all games concurrently
//   SPRL::UCTTree                                     uct/UCTTree.hpp:38-210: getDecisionNode / searchAndGetLeaves /
//                                                     evaluateAndBackpropLeaves / advanceDecision over a one-tree engine
//                                                     (the step-wise C ABI: sprl_begin_trees, sprl_search_batch,
//                                                     sprl_apply_evaluations, sprl_root_stats, sprl_advance)
//   SPRL::IAgent, UCTNetworkAgent, playGame           agents/*.hpp, evaluate/play.hpp:24-69: act / opponentAct drive the
//                                                     agent's own tree; the reference's Evaluate.cpp compiles unchanged and
//                                                     plays one game per playGame call on the GPU (playMatch is the batched form)
//
// What changes underneath: runIteration plays all `numGames` games CONCURRENTLY on one GPU
// (one warp per tree) instead of one after the other on a CPU core; the evaluator is a handle
// (uniform on device, or a traced network run on the GPU from device buffers) rather than an
// object whose evaluate() is called with host vectors; samples come back already embedded as
// the float planes runWorker writes (selfplay/GridWorker.hpp:146-171).
#ifndef SPRL_B200_VENEER_HPP
#define SPRL_B200_VENEER_HPP

#include "../../../../include/sprl_b200.h"

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <filesystem>
#include <iostream>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

namespace SPRL {

enum class Player : int8_t { NONE = -1, ZERO = 0, ONE = 1 };
enum class Piece : int8_t { NONE = -1, ZERO = 0, ONE = 1 };
using ActionIdx = int16_t;
using Value = float;
using SymmetryIdx = int8_t;

enum class InitQ { ZERO, PARENT, DROP_PARENT };

// games/GameActionDist.hpp:18-293: a fixed-size float vector with the reference's arithmetic -- sum() adds in index
// order, dividing by a scalar multiplies by its reciprocal (:283-293), pow / exp are libm's.
template <int ACTION_SIZE> struct GameActionDist : std::array<float, ACTION_SIZE> {
    static constexpr int SIZE = ACTION_SIZE;
    GameActionDist() { this->fill(0.0f); }
    float sum() const { float s = 0.0f; for (float v : *this) s += v; return s; }
    GameActionDist pow(float e) const { GameActionDist r; for (int i = 0; i < ACTION_SIZE; ++i) r[i] = std::pow((*this)[i], e); return r; }
    GameActionDist exp() const { GameActionDist r; for (int i = 0; i < ACTION_SIZE; ++i) r[i] = std::exp((*this)[i]); return r; }
    GameActionDist cumsum() const { GameActionDist r; float run = 0.0f; for (int i = 0; i < ACTION_SIZE; ++i) { run += (*this)[i]; r[i] = run; } return r; }
    GameActionDist operator/(float rhs) const { GameActionDist r; const float inv = 1.0f / rhs; for (int i = 0; i < ACTION_SIZE; ++i) r[i] = (*this)[i] * inv; return r; }
    GameActionDist operator*(float rhs) const { GameActionDist r; for (int i = 0; i < ACTION_SIZE; ++i) r[i] = (*this)[i] * rhs; return r; }
};
// games/GameNode.hpp:26-32, games/GridState.hpp:17-47
constexpr Player otherPlayer(Player player) { return player == Player::ZERO ? Player::ONE : (player == Player::ONE ? Player::ZERO : Player::NONE); }
constexpr Piece otherPiece(Piece piece) { return piece == Piece::ZERO ? Piece::ONE : (piece == Piece::ONE ? Piece::ZERO : Piece::NONE); }
constexpr Piece pieceFromPlayer(Player player) { return static_cast<Piece>(static_cast<int8_t>(player)); }
constexpr Player playerFromPiece(Piece piece) { return static_cast<Player>(static_cast<int8_t>(piece)); }

template <int BOARD_SIZE> using GridBoard = std::array<Piece, BOARD_SIZE>;

// games/GridState.hpp:56-115: the boards a network sees -- history[0] is the current board, higher indices go back in time,
// size() of them are valid -- and the player to move.
template <int BOARD_SIZE, int HISTORY_SIZE> class GridState {
public:
    static constexpr int BOARD = BOARD_SIZE, HISTORY = HISTORY_SIZE;
    GridState() { for (auto& b : m_history) b.fill(Piece::NONE); }
    GridState(std::array<GridBoard<BOARD_SIZE>, HISTORY_SIZE>&& history, int size, Player player)
        : m_history(std::move(history)), m_size(size), m_player(player) {}
    const std::array<GridBoard<BOARD_SIZE>, HISTORY_SIZE>& getHistory() const { return m_history; }
    int size() const { return m_size; }
    Player getPlayer() const { return m_player; }
private:
    std::array<GridBoard<BOARD_SIZE>, HISTORY_SIZE> m_history;
    int m_size { 0 };
    Player m_player { Player::NONE };
};

// errors of the C ABI as exceptions
class EngineError : public std::runtime_error {
public:
    EngineError(int code, const std::string& what) : std::runtime_error(what), code(code) {}
    int code;
};
inline void check(int rc) {
    if (rc != SPRL_OK) throw EngineError(rc, std::string("libsprl_b200: ") + sprl_last_error());
}
inline int currentDevice();
struct RawNode {};                   // constructs a node without asking the device for the start position

// games/GameNode.hpp:50-200 on the host: a node is a position the DEVICE computed.  The constructor asks for the start
// position, getAddChild(action) for the position after the line of actions from the root to this node plus `action`
// (sprl_env_line: one launch, the line is also Go's superko history) and caches the child like the reference; mask,
// player, winner, terminal flag, rewards and the GridState of the last HISTORY boards are the reference's accessors.
// An illegal action throws EngineError(SPRL_E_INVALID) where the reference asserts.
template <typename ImplNode, typename State, int ACTION_SIZE>
class GameNode {
public:
    using ActionDist = GameActionDist<ACTION_SIZE>;
    static constexpr int BOARD_CELLS = State::BOARD;

    GameNode() {
        int8_t player = 0, terminal = 0, winner = -1;
        std::array<int8_t, ACTION_SIZE> mask {};
        check(sprl_env_line(currentDevice(), ImplNode::GAME, 0, nullptr, m_cells.data(), &player, &terminal, &winner, mask.data()));
        fill(player, terminal, winner, mask.data());
    }
    explicit GameNode(RawNode) {}
    virtual ~GameNode() {}

    ImplNode* getParent() const { return m_parent; }

    ImplNode* getAddChild(ActionIdx action) {
        if (m_isTerminal) throw EngineError(SPRL_E_STATE, "getAddChild on a terminal node");
        if (action < 0 || action >= ACTION_SIZE || !(m_actionMask[action] > 0.0f))
            throw EngineError(SPRL_E_INVALID, "getAddChild: action " + std::to_string(action) + " is not legal here");
        if (!m_children[action]) {
            std::vector<int32_t> line;                                   // the actions from the root to the new child
            for (const GameNode* n = this; n->m_parent; n = n->m_parent) line.push_back(n->m_action);
            std::reverse(line.begin(), line.end());
            line.push_back(action);
            const size_t n = line.size() + 1;
            std::vector<int8_t> cells(n * BOARD_CELLS), player(n), terminal(n), winner(n), mask(n * ACTION_SIZE);
            check(sprl_env_line(currentDevice(), ImplNode::GAME, (int32_t)line.size(), line.data(), cells.data(), player.data(),
                                terminal.data(), winner.data(), mask.data()));
            auto child = std::make_unique<ImplNode>(RawNode {});
            child->m_parent = static_cast<ImplNode*>(this);
            child->m_action = action;
            std::copy(cells.end() - BOARD_CELLS, cells.end(), child->m_cells.begin());
            child->fill(player.back(), terminal.back(), winner.back(), mask.data() + (n - 1) * ACTION_SIZE);
            m_children[action] = std::move(child);
        }
        return m_children[action].get();
    }

    void pruneChildrenExcept(ActionIdx action) {
        for (ActionIdx i = 0; i < ACTION_SIZE; ++i)
            if (i != action) m_children[i] = nullptr;
    }

    Player getPlayer() const { return m_player; }
    Player getWinner() const { return m_winner; }
    bool isTerminal() const { return m_isTerminal; }
    const ActionDist& getActionMask() const { return m_actionMask; }

    State getGameState() const {
        std::array<GridBoard<State::BOARD>, State::HISTORY> history;
        int size = 0;
        for (const GameNode* n = this; n && size < State::HISTORY; n = n->m_parent, ++size)
            for (int i = 0; i < State::BOARD; ++i) history[size][i] = static_cast<Piece>(n->m_cells[i]);
        for (int t = size; t < State::HISTORY; ++t) history[t].fill(Piece::NONE);
        return State(std::move(history), size, m_player);
    }

    std::array<Value, 2> getRewards() const {              // the reference's getRewardsImpl of all three games
        if (m_winner == Player::ZERO) return { 1.0f, -1.0f };
        if (m_winner == Player::ONE) return { -1.0f, 1.0f };
        return { 0.0f, 0.0f };
    }

    std::string toString() const {
        std::string out;
        const int cols = ImplNode::COLS;
        for (int i = 0; i < State::BOARD; ++i) {
            out += m_cells[i] == 0 ? 'O' : (m_cells[i] == 1 ? 'X' : '.');
            if ((i + 1) % cols == 0) out += '\n';
        }
        return out;
    }

    const std::array<int8_t, State::BOARD>& cells() const { return m_cells; }     // Piece encoding, row-major

protected:
    void fill(int8_t player, int8_t terminal, int8_t winner, const int8_t* mask) {
        m_player = static_cast<Player>(player);
        m_isTerminal = terminal != 0;
        m_winner = static_cast<Player>(winner);
        for (int a = 0; a < ACTION_SIZE; ++a) m_actionMask[a] = mask[a] ? 1.0f : 0.0f;
    }
    ImplNode* m_parent { nullptr };
    std::array<std::unique_ptr<ImplNode>, ACTION_SIZE> m_children;
    ActionIdx m_action { 0 };
    ActionDist m_actionMask;
    Player m_player { Player::ZERO };
    Player m_winner { Player::NONE };
    bool m_isTerminal { false };
    std::array<int8_t, State::BOARD> m_cells {};
};

// ---- games: constants of games/*.hpp and tags selecting the device rules ----
constexpr int OTH_BOARD_WIDTH = 8;
constexpr int OTH_BOARD_SIZE = OTH_BOARD_WIDTH * OTH_BOARD_WIDTH;
constexpr int OTH_ACTION_SIZE = OTH_BOARD_SIZE + 1;
constexpr int OTH_HISTORY_SIZE = 1;
struct OthelloNode : GameNode<OthelloNode, GridState<OTH_BOARD_SIZE, OTH_HISTORY_SIZE>, OTH_ACTION_SIZE> {
    using State = GridState<OTH_BOARD_SIZE, OTH_HISTORY_SIZE>;
    using GameNode::GameNode;
    static constexpr int GAME = SPRL_GAME_OTHELLO, COLS = OTH_BOARD_WIDTH;
};

constexpr int C4_NUM_ROWS = 6;
constexpr int C4_NUM_COLS = 7;
constexpr int C4_BOARD_SIZE = C4_NUM_ROWS * C4_NUM_COLS;
constexpr int C4_ACTION_SIZE = C4_NUM_COLS;
constexpr int C4_HISTORY_SIZE = 1;
struct ConnectFourNode : GameNode<ConnectFourNode, GridState<C4_BOARD_SIZE, C4_HISTORY_SIZE>, C4_ACTION_SIZE> {
    using State = GridState<C4_BOARD_SIZE, C4_HISTORY_SIZE>;
    using GameNode::GameNode;
    static constexpr int GAME = SPRL_GAME_C4, COLS = C4_NUM_COLS;
};

#ifndef SPRL_GO_BOARD_WIDTH
#define SPRL_GO_BOARD_WIDTH 7                 // games/GoNode.hpp:16; 9 selects the 9x9 / komi 7.5 rules
#endif
constexpr int GO_BOARD_WIDTH = SPRL_GO_BOARD_WIDTH;
constexpr int GO_BOARD_SIZE = GO_BOARD_WIDTH * GO_BOARD_WIDTH;
constexpr int GO_ACTION_SIZE = GO_BOARD_SIZE + 1;
constexpr int GO_HISTORY_SIZE = 8;
struct GoNode : GameNode<GoNode, GridState<GO_BOARD_SIZE, GO_HISTORY_SIZE>, GO_ACTION_SIZE> {
    using State = GridState<GO_BOARD_SIZE, GO_HISTORY_SIZE>;
    using GameNode::GameNode;
    static constexpr int GAME = (GO_BOARD_WIDTH == 9) ? SPRL_GAME_GO9 : SPRL_GAME_GO7, COLS = GO_BOARD_WIDTH;
};

// ---- evaluators ----
// The reference's INetwork::evaluate(states, masks) is a host call per leaf batch.  Here an
// evaluator is a handle the engine queries for (a) which on-device evaluator to use and
// (b) for external networks, a forward function over device buffers.
template <typename State, int ACTION_SIZE>
class INetwork {
public:
    virtual ~INetwork() = default;
    using ActionDist = GameActionDist<ACTION_SIZE>;
    virtual int evaluatorKind() const = 0;                       // SPRL_EVAL_*
    // networks/INetwork.hpp:25-27 for host callers: prior over the legal actions and value of every state.  The engine never
    // calls this (its leaves are evaluated from device buffers, see forward()); evaluators that exist only on the device
    // (the test HashNetwork) do not offer it.
    virtual std::vector<std::pair<ActionDist, Value>> evaluate(const std::vector<State>& /*states*/, const std::vector<ActionDist>& /*masks*/) {
        throw EngineError(SPRL_E_STATE, "this evaluator has no host-side evaluate()");
    }
    // External networks only.  prepare(): allocate the device buffers of a leaf batch
    // (planes [batch, planes, rows, cols], logits [batch, actions], value [batch]) on `device`.
    // forward(): run the network on d_in and leave its outputs in d_logits / d_value, on `stream`.
    virtual int prepare(int /*device*/, int64_t /*batch*/, int /*planes*/, int /*rows*/, int /*cols*/, int /*actions*/,
                        float** /*d_in*/, float** /*d_logits*/, float** /*d_value*/) { return -1; }
    // attach(): what prepare() does without owning the buffers (second network of a match: it reads and writes
    // its half of the first network's buffers).
    virtual int attach(int /*device*/, int /*planes*/, int /*rows*/, int /*cols*/, int /*actions*/) { return -1; }
    virtual int forward(const float*, int64_t, float*, float*, void* /*cudaStream_t*/) { return -1; }
    // Device address of the number of rows in use (sprl_eval_rows); networks that can read their batch size on the
    // device skip the unused rows.
    virtual void setRowCount(const uint32_t* /*d_rows*/) {}
    // After a run: 0, or an SPRL_E_* code when the evaluator reported a failure (e.g. an activation outside the range
    // of the library evaluator's fp16 split); the message is in sprl_last_error().
    virtual int checkStatus() { return 0; }
    virtual int getNumEvals() { return (int)m_numEvals; }
    void addEvals(uint64_t n) { m_numEvals += n; }
protected:
    uint64_t m_numEvals { 0 };
};

template <typename State, int ACTION_SIZE>
class RandomNetwork : public INetwork<State, ACTION_SIZE> {       // networks/RandomNetwork.hpp
public:
    using ActionDist = GameActionDist<ACTION_SIZE>;
    int evaluatorKind() const override { return SPRL_EVAL_UNIFORM; }
    // networks/RandomNetwork.hpp:21-49: uniform over the legal actions, value 0
    std::vector<std::pair<ActionDist, Value>> evaluate(const std::vector<State>& states, const std::vector<ActionDist>& masks) override {
        std::vector<std::pair<ActionDist, Value>> out;
        out.reserve(states.size());
        for (size_t b = 0; b < states.size(); ++b) out.emplace_back(masks[b] / masks[b].sum(), 0.0f);
        this->addEvals(states.size());
        return out;
    }
};

// Deterministic test evaluator (parity runs), evaluated on the device.
template <typename State, int ACTION_SIZE>
class HashNetwork : public INetwork<State, ACTION_SIZE> {
public:
    int evaluatorKind() const override { return SPRL_EVAL_HASHNET; }
};

// ---- symmetrizers (symmetry/ISymmetrizer.hpp:17-57) ----
// Inside the engine the PRESENCE of a symmetrizer selects the game's symmetry group on the device (leaf encoding, inverse
// mapping of the policies, the symmetric samples).  The host functions below are the reference's interface for callers that
// hold a GridState or an action distribution themselves; they use the same maps as the device (new[f_s(i)] = old[i]) and
// tests/hostcheck/gamenode.cpp checks them against the device's symmetric samples.
template <typename State, int ACTION_SIZE>
class ISymmetrizer {
public:
    using ActionDist = GameActionDist<ACTION_SIZE>;
    virtual ~ISymmetrizer() = default;
    virtual int numSymmetries() const = 0;
    virtual SymmetryIdx inverseSymmetry(SymmetryIdx symmetry) const = 0;
    virtual std::vector<State> symmetrizeState(const State& state, const std::vector<SymmetryIdx>& symmetries) const = 0;
    virtual std::vector<ActionDist> symmetrizeActionDist(const ActionDist& actionDist, const std::vector<SymmetryIdx>& symmetries) const = 0;
};

// the board maps shared by a symmetrizer's two functions: cellTo(s, i) = f_s(i), actionTo(s, a)
template <typename Derived, typename State, int ACTION_SIZE>
class GridSymmetrizerBase : public ISymmetrizer<State, ACTION_SIZE> {
public:
    using ActionDist = GameActionDist<ACTION_SIZE>;
    std::vector<State> symmetrizeState(const State& state, const std::vector<SymmetryIdx>& symmetries) const override {
        std::vector<State> out;
        out.reserve(symmetries.size());
        for (SymmetryIdx s : symmetries) {
            std::array<GridBoard<State::BOARD>, State::HISTORY> history;
            for (int t = 0; t < State::HISTORY; ++t) {
                history[t].fill(Piece::NONE);
                if (t < state.size())
                    for (int i = 0; i < State::BOARD; ++i) history[t][Derived::cellTo(s, i)] = state.getHistory()[t][i];
            }
            out.emplace_back(std::move(history), state.size(), state.getPlayer());
        }
        return out;
    }
    std::vector<ActionDist> symmetrizeActionDist(const ActionDist& actionDist, const std::vector<SymmetryIdx>& symmetries) const override {
        std::vector<ActionDist> out;
        out.reserve(symmetries.size());
        for (SymmetryIdx s : symmetries) {
            ActionDist d;
            for (int a = 0; a < ACTION_SIZE; ++a) d[Derived::actionTo(s, a)] = actionDist[a];
            out.push_back(d);
        }
        return out;
    }
};

template <int BOARD_WIDTH, int HISTORY_SIZE>
class D4GridSymmetrizer
    : public GridSymmetrizerBase<D4GridSymmetrizer<BOARD_WIDTH, HISTORY_SIZE>, GridState<BOARD_WIDTH * BOARD_WIDTH, HISTORY_SIZE>, BOARD_WIDTH * BOARD_WIDTH + 1> {
public:
    int numSymmetries() const override { return 8; }
    SymmetryIdx inverseSymmetry(SymmetryIdx s) const override {
        static constexpr SymmetryIdx inv[8] = { 0, 3, 2, 1, 4, 5, 6, 7 };     // symmetry/D4GridSymmetrizer.hpp:47-50
        return inv[s];
    }
    // symmetry/D4GridSymmetrizer.hpp:106-117 as one table: (r, c) goes to
    // 0 (r,c)  1 (c,w-1-r)  2 (w-1-r,w-1-c)  3 (w-1-c,r)  4 (r,w-1-c)  5 (w-1-c,w-1-r)  6 (w-1-r,c)  7 (c,r)
    static int cellTo(SymmetryIdx s, int from) {
        const int w = BOARD_WIDTH, r = from / w, c = from % w, x = w - 1;
        switch (s) {
        case 1: return c * w + (x - r);
        case 2: return (x - r) * w + (x - c);
        case 3: return (x - c) * w + r;
        case 4: return r * w + (x - c);
        case 5: return (x - c) * w + (x - r);
        case 6: return (x - r) * w + c;
        case 7: return c * w + r;
        default: return from;
        }
    }
    static int actionTo(SymmetryIdx s, int a) { return a == BOARD_WIDTH * BOARD_WIDTH ? a : cellTo(s, a); }     // the pass stays
};

class ConnectFourSymmetrizer : public GridSymmetrizerBase<ConnectFourSymmetrizer, GridState<C4_BOARD_SIZE, C4_HISTORY_SIZE>, C4_ACTION_SIZE> {
public:
    int numSymmetries() const override { return 2; }
    SymmetryIdx inverseSymmetry(SymmetryIdx s) const override { return s; }
    static int cellTo(SymmetryIdx s, int from) { return s == 1 ? (from / C4_NUM_COLS) * C4_NUM_COLS + (C4_NUM_COLS - 1 - from % C4_NUM_COLS) : from; }
    static int actionTo(SymmetryIdx s, int a) { return s == 1 ? C4_NUM_COLS - 1 - a : a; }
};

// ---- engine handle -------------------------------------------------------------------------

// Run-wide knobs a reference main cannot express (it has no notion of a device).
struct DeviceOptions {
    int device = 0;
    uint64_t seed = 0;
    int numSlots = 0;          // 0 = one slot per game
    uint64_t firstGame = 0;    // stream id of the first game of the next runIteration
    uint64_t gameStride = 1;
    bool fixSymmetryMask = false;   // not the reference: symmetrise the legal mask with the state (repairs quirk Q3)
    int64_t treeUnits = 1 << 20;    // slab size (16-byte units) of a UCTTree object's device tree; 0 = sized for 2^30 descents
};
// The reference seeds its generator from std::random_device (constants.hpp:4 SEED = 0, utils/random.cpp:38-47), so two
// runs never replay the same games; same here unless SPRL_SEED pins the run seed (tests, reproducible deployments).
inline DeviceOptions& deviceOptions() {
    static DeviceOptions o = [] {
        DeviceOptions d;
        const char* env = std::getenv("SPRL_SEED");
        if (env) d.seed = std::strtoull(env, nullptr, 10);
        else {
            std::random_device rd;
            d.seed = ((uint64_t)rd() << 32) | (uint64_t)rd();
            std::cout << "Run seed " << d.seed << " (std::random_device; set SPRL_SEED to pin it)" << std::endl;
        }
        return d;
    }();
    return o;
}
inline int currentDevice() { return deviceOptions().device; }

inline int initQCode(InitQ q) {
    return q == InitQ::PARENT ? SPRL_INITQ_PARENT : (q == InitQ::DROP_PARENT ? SPRL_INITQ_DROP_PARENT : SPRL_INITQ_ZERO);
}

// ---- runIteration (selfplay/SelfPlay.hpp:203-248) --------------------------------------------
// Returns the embedded samples: states [n, 2H+1, R, C], distributions [n, A], outcomes [n].
template <typename ImplNode, typename State, int ACTION_SIZE>
std::tuple<std::vector<float>, std::vector<float>, std::vector<Value>>
runIteration(INetwork<State, ACTION_SIZE>* network, int numGames,
             int numTraversals, int maxBatchSize, int maxQueueSize,
             float dirEps, float dirAlpha, InitQ initQMethod,
             ISymmetrizer<State, ACTION_SIZE>* symmetrizer, bool addNoise = true) {
    const DeviceOptions& opt = deviceOptions();
    sprl_config cfg;
    check(sprl_default_config(ImplNode::GAME, &cfg));
    cfg.device = opt.device; cfg.seed = opt.seed;
    cfg.evaluator = network->evaluatorKind();
    cfg.num_slots = opt.numSlots > 0 ? std::min(opt.numSlots, numGames) : numGames;
    cfg.max_games = numGames;
    cfg.sims = numTraversals; cfg.max_batch = maxBatchSize; cfg.max_queue = maxQueueSize;
    cfg.dir_eps = dirEps; cfg.dir_alpha = dirAlpha;
    cfg.init_q = initQCode(initQMethod);
    cfg.add_noise = addNoise ? 1 : 0;
    cfg.use_sym = symmetrizer != nullptr ? 1 : 0;
    cfg.fix_symmetry_mask = opt.fixSymmetryMask ? 1 : 0;
    sprl_engine* e = nullptr;
    check(sprl_create(&cfg, &e));
    struct Guard { sprl_engine* e; ~Guard() { sprl_destroy(e); } } guard { e };
    check(sprl_set_game_stride(e, opt.gameStride));

    sprl_game_info gi;
    check(sprl_game_info_get(ImplNode::GAME, &gi));
    struct Ctx { INetwork<State, ACTION_SIZE>* net; } ctx { network };
    sprl_forward_fn fwd = nullptr;
    if (cfg.evaluator == SPRL_EVAL_EXTERNAL) {
        fwd = [](void* user, const float* in, int64_t batch, float* logits, float* value, void* stream) -> int {
            return static_cast<Ctx*>(user)->net->forward(in, batch, logits, value, stream);
        };
        // the evaluator buffers are owned by the network object
        float *d_in = nullptr, *d_logits = nullptr, *d_value = nullptr;
        if (network->prepare(cfg.device, sprl_eval_batch(e), 2 * gi.history + 1, gi.rows, gi.cols, gi.actions, &d_in, &d_logits, &d_value) != 0)
            throw EngineError(SPRL_E_STATE, "network could not allocate its device buffers");
        check(sprl_bind_eval_buffers(e, d_in, d_logits, d_value));
        const uint32_t* d_rows = nullptr;
        check(sprl_eval_rows(e, &d_rows));
        network->setRowCount(d_rows);
    }
    check(sprl_run_iteration(e, opt.firstGame, numGames, fwd, &ctx));
    check(network->checkStatus());

    int64_t nMoves = 0, nSamples = 0;
    check(sprl_iteration_counts(e, &nMoves, &nSamples));
    const size_t row = (size_t)(2 * gi.history + 1) * gi.cells;
    std::vector<float> states((size_t)nSamples * row), dists((size_t)nSamples * gi.actions);
    std::vector<Value> outcomes((size_t)nSamples);
    int64_t got = 0;
    check(sprl_collect_samples(e, nSamples, states.data(), dists.data(), outcomes.data(), &got));
    sprl_stats st;
    check(sprl_get_stats(e, &st));
    network->addEvals(st.evals);
    std::cout << numGames << " games played, " << nSamples << " states collected.\n";
    return { std::move(states), std::move(dists), std::move(outcomes) };
}

// ---- selfPlay (selfplay/SelfPlay.hpp:50-192) ---------------------------------------------------
// ONE game from the start position in the reference's own types: the symmetrised GridStates over time, the symmetrised
// action distributions UCT produced and the outcome for the player to move at each of them -- decoded from the embedded
// planes of a one-game runIteration (plane 2t / 2t+1 = the mover's / the opponent's stones t plies back, last plane =
// "player ZERO to move", selfplay/GridWorker.hpp:146-171).  Successive calls draw from successive game streams.
template <typename ImplNode, typename State, int ACTION_SIZE>
std::tuple<std::vector<State>, std::vector<GameActionDist<ACTION_SIZE>>, std::vector<Value>>
selfPlay(std::unique_ptr<GameNode<ImplNode, State, ACTION_SIZE>> rootNode,
         INetwork<State, ACTION_SIZE>* network,
         int numTraversals, int maxBatchSize, int maxQueueSize,
         float dirEps, float dirAlpha, InitQ initQMethod,
         ISymmetrizer<State, ACTION_SIZE>* symmetrizer, bool addNoise = true) {
    if (!rootNode || rootNode->getParent() != nullptr)
        throw EngineError(SPRL_E_INVALID, "selfPlay: the engine plays from the start position (pass a fresh root node)");
    DeviceOptions& opt = deviceOptions();
    auto [planes, flat, outcomes] = runIteration<ImplNode, State, ACTION_SIZE>(network, 1, numTraversals, maxBatchSize, maxQueueSize,
                                                                              dirEps, dirAlpha, initQMethod, symmetrizer, addNoise);
    opt.firstGame += opt.gameStride;
    const int nsym = symmetrizer ? symmetrizer->numSymmetries() : 1;
    const size_t n = outcomes.size(), row = (size_t)(2 * State::HISTORY + 1) * State::BOARD;
    std::vector<State> states;
    std::vector<GameActionDist<ACTION_SIZE>> distributions(n);
    states.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        const float* p = planes.data() + i * row;
        const Player player = p[(size_t)2 * State::HISTORY * State::BOARD] > 0.5f ? Player::ZERO : Player::ONE;
        const Piece mine = pieceFromPlayer(player), theirs = otherPiece(mine);
        const int size = (int)std::min<size_t>(i / nsym + 1, State::HISTORY);
        std::array<GridBoard<State::BOARD>, State::HISTORY> history;
        for (int t = 0; t < State::HISTORY; ++t)
            for (int c = 0; c < State::BOARD; ++c)
                history[t][c] = (t < size && p[(size_t)(2 * t) * State::BOARD + c] > 0.5f) ? mine
                              : ((t < size && p[(size_t)(2 * t + 1) * State::BOARD + c] > 0.5f) ? theirs : Piece::NONE);
        states.emplace_back(std::move(history), size, player);
        for (int a = 0; a < ACTION_SIZE; ++a) distributions[i][a] = flat[i * ACTION_SIZE + a];
    }
    return { std::move(states), std::move(distributions), std::move(outcomes) };
}

// ---- playMatch: the loop of Evaluate.cpp:93-157 ------------------------------------------------
// `numGames` games between two networks, each side with its own tree, symmetrizer and init-Q; game t gives
// Player ZERO to network t % 2; the side to move searches numTraversals descents and plays its most-visited
// action (agents/UCTNetworkAgent.hpp:45-105), the other tree follows (opponentAct, :107-109).  Trees are built as
// Evaluate.cpp builds them: Dirichlet noise on, eps 0.25, alpha 0.1, default uWeight 1.0.
struct MatchResult { int64_t wins0, wins1, draws; std::vector<int8_t> winners; std::vector<int32_t> moves; };

struct MatchSearch { float dirEps = 0.25f, dirAlpha = 0.1f, uWeight = 1.0f; bool addNoise = true; };     // Evaluate.cpp:95-111

template <typename ImplNode, typename State, int ACTION_SIZE>
MatchResult runMatch(INetwork<State, ACTION_SIZE>* network0, INetwork<State, ACTION_SIZE>* network1, int numGames,
                     int numTraversals, int maxBatchSize, int maxQueueSize,
                     ISymmetrizer<State, ACTION_SIZE>* symmetrizer0, InitQ initQ0,
                     ISymmetrizer<State, ACTION_SIZE>* symmetrizer1, InitQ initQ1,
                     const MatchSearch& search, uint64_t firstGame, uint64_t gameStride) {
    const DeviceOptions& opt = deviceOptions();
    INetwork<State, ACTION_SIZE>* nets[2] = { network0, network1 };
    sprl_agent_config agents[2] = {
        { network0->evaluatorKind(), symmetrizer0 != nullptr ? 1 : 0, initQCode(initQ0), 0 },
        { network1->evaluatorKind(), symmetrizer1 != nullptr ? 1 : 0, initQCode(initQ1), 0 } };
    const bool external = agents[0].evaluator == SPRL_EVAL_EXTERNAL || agents[1].evaluator == SPRL_EVAL_EXTERNAL;
    sprl_config cfg;
    check(sprl_default_config(ImplNode::GAME, &cfg));
    cfg.device = opt.device; cfg.seed = opt.seed;
    cfg.evaluator = external ? SPRL_EVAL_EXTERNAL : SPRL_EVAL_UNIFORM;
    const int pairs = opt.numSlots > 0 ? std::min(opt.numSlots, numGames) : numGames;
    cfg.num_slots = 2 * pairs;
    cfg.max_games = numGames;
    cfg.sims = numTraversals; cfg.max_batch = maxBatchSize; cfg.max_queue = maxQueueSize;
    cfg.dir_eps = search.dirEps; cfg.dir_alpha = search.dirAlpha; cfg.u_weight = search.uWeight; cfg.add_noise = search.addNoise ? 1 : 0;
    cfg.fix_symmetry_mask = opt.fixSymmetryMask ? 1 : 0;
    sprl_engine* e = nullptr;
    check(sprl_create(&cfg, &e));
    struct Guard { sprl_engine* e; ~Guard() { sprl_destroy(e); } } guard { e };
    check(sprl_set_game_stride(e, gameStride));
    sprl_game_info gi;
    check(sprl_game_info_get(ImplNode::GAME, &gi));

    struct Ctx { INetwork<State, ACTION_SIZE>* net[2]; int64_t half, rowIn, rowLogits; } ctx { { nullptr, nullptr }, 0, 0, 0 };
    ctx.half = (int64_t)pairs * maxQueueSize;
    ctx.rowIn = (int64_t)(2 * gi.history + 1) * gi.cells;
    ctx.rowLogits = gi.actions;
    if (external) {
        // the first external network owns the buffers of the whole batch; each network serves its own half
        float *d_in = nullptr, *d_logits = nullptr, *d_value = nullptr;
        const int owner = agents[0].evaluator == SPRL_EVAL_EXTERNAL ? 0 : 1;
        if (nets[owner]->prepare(cfg.device, sprl_eval_batch(e), 2 * gi.history + 1, gi.rows, gi.cols, gi.actions, &d_in, &d_logits, &d_value) != 0)
            throw EngineError(SPRL_E_STATE, "network could not allocate its device buffers");
        ctx.net[owner] = nets[owner];
        if (owner == 0 && agents[1].evaluator == SPRL_EVAL_EXTERNAL) {
            if (nets[1] != nets[0] && nets[1]->attach(cfg.device, 2 * gi.history + 1, gi.rows, gi.cols, gi.actions) != 0)
                throw EngineError(SPRL_E_STATE, "second network could not be attached to the device");
            ctx.net[1] = nets[1];
        }
        check(sprl_bind_eval_buffers(e, d_in, d_logits, d_value));
        const uint32_t* d_rows = nullptr;
        check(sprl_eval_rows(e, &d_rows));
        // each network evaluates the rows its side queued (count read on the device); ONE network object serving both
        // sides has a single row-count slot, so it evaluates both halves at full size instead
        for (int k = 0; k < 2; ++k)
            if (ctx.net[k]) ctx.net[k]->setRowCount(ctx.net[0] == ctx.net[1] ? nullptr : d_rows + k);
    }
    sprl_forward_fn fwd = [](void* user, const float* in, int64_t, float* logits, float* value, void* stream) -> int {
        Ctx* c = static_cast<Ctx*>(user);
        for (int k = 0; k < 2; ++k) {
            if (!c->net[k]) continue;
            int rc = c->net[k]->forward(in + k * c->half * c->rowIn, c->half, logits + k * c->half * c->rowLogits,
                                        value + k * c->half, stream);
            if (rc) return rc;
        }
        return 0;
    };
    check(sprl_run_match(e, agents, firstGame, numGames, external ? fwd : nullptr, &ctx));
    for (int k = 0; k < 2; ++k) check(nets[k]->checkStatus());
    MatchResult r { 0, 0, 0, std::vector<int8_t>((size_t)numGames), std::vector<int32_t>((size_t)numGames) };
    int64_t wins[2] = { 0, 0 };
    check(sprl_match_results(e, numGames, r.winners.data(), r.moves.data(), nullptr, wins, &r.draws));
    r.wins0 = wins[0]; r.wins1 = wins[1];
    return r;
}

template <typename ImplNode, typename State, int ACTION_SIZE>
MatchResult playMatch(INetwork<State, ACTION_SIZE>* network0, INetwork<State, ACTION_SIZE>* network1, int numGames,
                      int numTraversals, int maxBatchSize, int maxQueueSize,
                      ISymmetrizer<State, ACTION_SIZE>* symmetrizer0, InitQ initQ0,
                      ISymmetrizer<State, ACTION_SIZE>* symmetrizer1, InitQ initQ1) {
    return runMatch<ImplNode, State, ACTION_SIZE>(network0, network1, numGames, numTraversals, maxBatchSize, maxQueueSize,
                                                  symmetrizer0, initQ0, symmetrizer1, initQ1, MatchSearch {},
                                                  deviceOptions().firstGame, deviceOptions().gameStride);
}

// ---- UCTTree / IAgent / UCTNetworkAgent / playGame as the reference spells them ------------------------
// uct/UCTTree.hpp:38-210 over ONE device tree (a one-slot engine created at the first search, when the batch / queue
// sizes and the evaluator are known).  The nodes live in device slabs; getDecisionNode() returns a host snapshot of the
// decision node: its EdgeStatistics (uct/UCTNode.hpp:45-60), N() / W(), player, terminal flag, winner and legal mask.
// Random draws come from the tree's own stream (deviceOptions().seed, deviceOptions().firstGame + trees made so far).
template <typename ImplNode, typename State, int ACTION_SIZE>
class UCTTree {
public:
    using ActionDist = GameActionDist<ACTION_SIZE>;
    struct EdgeStatistics { ActionDist m_childPriors, m_totalValues, m_numVisits; };
    struct DecisionNode {
        EdgeStatistics stats;
        ActionDist mask;
        float n = 0.0f, w = 0.0f;
        Player player = Player::ZERO, winner = Player::NONE;
        bool terminal = false;
        const EdgeStatistics* getEdgeStatistics() const { return &stats; }
        const ActionDist& getActionMask() const { return mask; }
        float N() const { return n; }
        float W() const { return w; }
        Player getPlayer() const { return player; }
        Player getWinner() const { return winner; }
        bool isTerminal() const { return terminal; }
        std::array<Value, 2> getRewards() const {                         // games/GameNode.hpp:170-186 through the winner
            if (winner == Player::ZERO) return { 1.0f, -1.0f };
            if (winner == Player::ONE) return { -1.0f, 1.0f };
            return { 0.0f, 0.0f };
        }
    };
    struct Leaf { int index; };       // a queued leaf of the last searchAndGetLeaves; it lives on the device

    UCTTree(std::unique_ptr<ImplNode> gameRoot, float dirEps, float dirAlpha, InitQ initQMethod,
            ISymmetrizer<State, ACTION_SIZE>* symmetrizer, bool addNoise = true)
        : dirEps(dirEps), dirAlpha(dirAlpha), initQMethod(initQMethod), symmetrizer(symmetrizer), addNoise(addNoise) {
        (void)gameRoot;
        static uint64_t treesMade = 0;
        streamId = deviceOptions().firstGame + deviceOptions().gameStride * treesMade++;
    }
    UCTTree(const UCTTree&) = delete;
    UCTTree& operator=(const UCTTree&) = delete;
    ~UCTTree() { if (engine) sprl_destroy(engine); }

    // uct/UCTTree.hpp:62
    const DecisionNode* getDecisionNode() {
        ensureAny();
        if (!snapshotValid) refresh();
        return &snapshot;
    }

    // uct/UCTTree.hpp:76-114: one batch of up to maxBatchSize descents that stops early at maxQueueSize queued leaves;
    // terminal and cached leaves are backed up at once.  Returns the queued leaves and the number of descents.
    std::pair<std::vector<Leaf>, int> searchAndGetLeaves(int maxBatchSize, int maxQueueSize, INetwork<State, ACTION_SIZE>* network,
                                                         float uWeight = 1.0f) {
        ensureEngine(maxBatchSize, maxQueueSize, network, uWeight);
        int32_t before = 0, after = 0, queued = 0;
        check(sprl_root_stats(engine, 1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &before, nullptr, &queued));
        if (queued != 0) throw EngineError(SPRL_E_STATE, "searchAndGetLeaves: the previous batch's leaves were not passed to evaluateAndBackpropLeaves");
        check(sprl_search_batch(engine));
        check(sprl_root_stats(engine, 1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &after, nullptr, &queued));
        snapshotValid = false;
        std::vector<Leaf> leaves;
        for (int i = 0; i < queued; ++i) leaves.push_back(Leaf { i });
        return { leaves, (int)(after - before) };
    }

    // uct/UCTTree.hpp:124-184: evaluate the queued leaves (one random symmetry each), expand and back up
    void evaluateAndBackpropLeaves(const std::vector<Leaf>& leaves, INetwork<State, ACTION_SIZE>* network) {
        if (!engine) throw EngineError(SPRL_E_STATE, "evaluateAndBackpropLeaves before searchAndGetLeaves");
        if (network != net) throw EngineError(SPRL_E_INVALID, "a tree keeps the evaluator of its first search");
        if (leaves.empty()) return;
        if (net->evaluatorKind() == SPRL_EVAL_EXTERNAL) {
            if (net->forward(dIn, sprl_eval_batch(engine), dLogits, dValue, nullptr) != 0)
                throw EngineError(SPRL_E_STATE, "the network's forward failed");
        }
        check(sprl_apply_evaluations(engine));
        net->addEvals((uint64_t)leaves.size());
        snapshotValid = false;
    }

    // uct/UCTTree.hpp:197-210
    void advanceDecision(ActionIdx action) {
        ensureAny();
        const int32_t a = (int32_t)action;
        check(sprl_advance(engine, &a, 1));
        history.push_back(a);
        snapshotValid = false;
    }

    float dirEps, dirAlpha;
    InitQ initQMethod;
    ISymmetrizer<State, ACTION_SIZE>* symmetrizer;
    bool addNoise;
    uint64_t streamId;

private:
    // The device tree needs the batch / queue sizes and the evaluator, which the reference only passes to the first
    // searchAndGetLeaves.  A tree that is inspected or advanced before its first search (the `while (!isTerminal())` of
    // every move loop; opponentAct when the other side moves first) lives in a provisional engine; the first search
    // replaces it and replays the advances -- an unsearched tree has drawn nothing from its stream, so nothing is lost.
    void ensureAny() { if (!engine) createEngine(1, 1, nullptr, 1.0f); }
    void ensureEngine(int maxBatchSize, int maxQueueSize, INetwork<State, ACTION_SIZE>* network, float uWeight) {
        if (engine && !provisional) {
            if (network != net || maxBatchSize != batch || maxQueueSize != queue || uWeight != uw)
                throw EngineError(SPRL_E_INVALID, "a tree keeps the batch size, queue size, evaluator and uWeight of its first search");
            return;
        }
        if (engine) { sprl_destroy(engine); engine = nullptr; }
        createEngine(maxBatchSize, maxQueueSize, network, uWeight);
    }
    void createEngine(int maxBatchSize, int maxQueueSize, INetwork<State, ACTION_SIZE>* network, float uWeight) {
        const DeviceOptions& opt = deviceOptions();
        sprl_config cfg;
        check(sprl_default_config(ImplNode::GAME, &cfg));
        cfg.device = opt.device; cfg.seed = opt.seed;
        cfg.evaluator = network ? network->evaluatorKind() : SPRL_EVAL_UNIFORM;
        cfg.num_slots = 1; cfg.max_games = 1;
        cfg.sims = 1 << 30;                              // budgets are the caller's: sprl_search_batch runs one batch per call
        cfg.max_batch = maxBatchSize; cfg.max_queue = maxQueueSize;
        cfg.dir_eps = dirEps; cfg.dir_alpha = dirAlpha; cfg.u_weight = uWeight;
        cfg.init_q = initQCode(initQMethod);
        cfg.add_noise = addNoise ? 1 : 0;
        cfg.use_sym = symmetrizer != nullptr ? 1 : 0;
        cfg.fix_symmetry_mask = opt.fixSymmetryMask ? 1 : 0;
        cfg.units_per_tree = opt.treeUnits;
        check(sprl_create(&cfg, &engine));
        sprl_game_info gi;
        check(sprl_game_info_get(ImplNode::GAME, &gi));
        if (cfg.evaluator == SPRL_EVAL_EXTERNAL) {
            if (network->prepare(cfg.device, sprl_eval_batch(engine), 2 * gi.history + 1, gi.rows, gi.cols, gi.actions, &dIn, &dLogits, &dValue) != 0)
                throw EngineError(SPRL_E_STATE, "network could not allocate its device buffers");
            check(sprl_bind_eval_buffers(engine, dIn, dLogits, dValue));
            const uint32_t* d_rows = nullptr;
            check(sprl_eval_rows(engine, &d_rows));
            network->setRowCount(d_rows);
        }
        check(sprl_begin_trees(engine, streamId, 1));
        for (int32_t a : history) check(sprl_advance(engine, &a, 1));
        net = network; batch = maxBatchSize; queue = maxQueueSize; uw = uWeight;
        provisional = network == nullptr;
        snapshotValid = false;
    }
    void refresh() {
        std::array<float, ACTION_SIZE> n {}, w {}, p {};
        std::array<int8_t, ACTION_SIZE> m {};
        int8_t player = 0, terminal = 0, winner = -1;
        check(sprl_root_stats(engine, 1, n.data(), w.data(), p.data(), &snapshot.n, &snapshot.w, &player, &terminal, &winner, nullptr, m.data(), nullptr));
        for (int i = 0; i < ACTION_SIZE; ++i) {
            snapshot.stats.m_numVisits[i] = n[i]; snapshot.stats.m_totalValues[i] = w[i]; snapshot.stats.m_childPriors[i] = p[i];
            snapshot.mask[i] = m[i] ? 1.0f : 0.0f;
        }
        snapshot.player = static_cast<Player>(player);
        snapshot.winner = static_cast<Player>(winner);
        snapshot.terminal = terminal != 0;
        snapshotValid = true;
    }
    sprl_engine* engine = nullptr;
    INetwork<State, ACTION_SIZE>* net = nullptr;
    int batch = 0, queue = 0;
    float uw = 1.0f;
    bool provisional = false, snapshotValid = false;
    std::vector<int32_t> history;       // actions advanced so far
    float *dIn = nullptr, *dLogits = nullptr, *dValue = nullptr;
    DecisionNode snapshot;
};

template <typename ImplNode, typename State, int ACTION_SIZE>
class IAgent {                                                       // agents/IAgent.hpp:16-38
public:
    using ActionDist = GameActionDist<ACTION_SIZE>;
    virtual ~IAgent() = default;
    virtual ActionIdx act(const GameNode<ImplNode, State, ACTION_SIZE>* gameNode, bool verbose = false) const = 0;
    virtual void opponentAct(const ActionIdx action) const = 0;
};

template <typename ImplNode, typename State, int ACTION_SIZE>
class UCTNetworkAgent : public IAgent<ImplNode, State, ACTION_SIZE> { // agents/UCTNetworkAgent.hpp:21-108
public:
    UCTNetworkAgent(INetwork<State, ACTION_SIZE>* network, UCTTree<ImplNode, State, ACTION_SIZE>* tree,
                    int numTraversals, int maxBatchSize, int maxQueueSize)
        : network(network), tree(tree), numTraversals(numTraversals), maxBatchSize(maxBatchSize), maxQueueSize(maxQueueSize) {}
    // :42-105: search numTraversals descents in the agent's own tree, play the FIRST most-visited action, advance
    ActionIdx act(const GameNode<ImplNode, State, ACTION_SIZE>* /*gameNode*/, bool /*verbose*/ = false) const override {
        int traversals = 0;
        while (traversals < numTraversals) {
            auto [leaves, n] = tree->searchAndGetLeaves(maxBatchSize, maxQueueSize, network);
            if (!leaves.empty()) tree->evaluateAndBackpropLeaves(leaves, network);
            traversals += n;
        }
        const auto& visits = tree->getDecisionNode()->getEdgeStatistics()->m_numVisits;
        const ActionIdx action = (ActionIdx)std::distance(visits.begin(), std::max_element(visits.begin(), visits.end()));
        tree->advanceDecision(action);
        return action;
    }
    // :106-108
    void opponentAct(const ActionIdx action) const override { tree->advanceDecision(action); }
    INetwork<State, ACTION_SIZE>* network;
    UCTTree<ImplNode, State, ACTION_SIZE>* tree;
    int numTraversals, maxBatchSize, maxQueueSize;
};

// evaluate/play.hpp:24-69: one game from the start position between agents[0] (Player ZERO) and agents[1]; returns the
// winner.  Both agents must be UCTNetworkAgents with the same search budget and noise parameters (as Evaluate.cpp builds
// them); every call plays a fresh game stream.
template <typename ImplNode, typename State, int ACTION_SIZE>
Player playGame(GameNode<ImplNode, State, ACTION_SIZE>* rootNode, std::array<IAgent<ImplNode, State, ACTION_SIZE>*, 2> agents,
                bool verbose = false) {
    (void)rootNode; (void)verbose;
    using Agent = UCTNetworkAgent<ImplNode, State, ACTION_SIZE>;
    Agent* a0 = dynamic_cast<Agent*>(agents[0]);
    Agent* a1 = dynamic_cast<Agent*>(agents[1]);
    if (!a0 || !a1) throw EngineError(SPRL_E_INVALID, "playGame: both agents must be UCTNetworkAgents (interactive agents are out of scope)");
    if (a0->numTraversals != a1->numTraversals || a0->maxBatchSize != a1->maxBatchSize || a0->maxQueueSize != a1->maxQueueSize ||
        a0->tree->dirEps != a1->tree->dirEps || a0->tree->dirAlpha != a1->tree->dirAlpha || a0->tree->addNoise != a1->tree->addNoise)
        throw EngineError(SPRL_E_INVALID, "playGame: the two agents must share the search budget and the noise parameters");
    static uint64_t gamesPlayed = 0;
    MatchSearch search;
    search.dirEps = a0->tree->dirEps; search.dirAlpha = a0->tree->dirAlpha; search.addNoise = a0->tree->addNoise;
    // an even game id gives Player ZERO to the first agent (sprl_match_begin)
    MatchResult r = runMatch<ImplNode, State, ACTION_SIZE>(a0->network, a1->network, 1, a0->numTraversals, a0->maxBatchSize,
                                                           a0->maxQueueSize, a0->tree->symmetrizer, a0->tree->initQMethod,
                                                           a1->tree->symmetrizer, a1->tree->initQMethod, search,
                                                           deviceOptions().firstGame + 2 * gamesPlayed++, 1);
    return static_cast<Player>(r.winners[0]);
}

// ---- waitModelPath / runWorker (selfplay/GridWorker.hpp:35-55,84-198) ------------------------
constexpr int MODEL_PATH_WAIT_INTERVAL = 30;

inline std::string waitModelPath(int iteration, const std::string& runName) {
    if (iteration == -1) return "random";
    std::string modelPath;
    do {
        modelPath = "data/models/" + runName + "/traced_" + runName + "_iteration_" + std::to_string(iteration) + ".pt";
        if (!std::filesystem::exists(modelPath)) {
            std::cout << "Spinning on traced model from iteration " << iteration << "..." << std::endl;
            std::this_thread::sleep_for(std::chrono::seconds(MODEL_PATH_WAIT_INTERVAL));
        }
    } while (!std::filesystem::exists(modelPath));
    std::this_thread::sleep_for(std::chrono::seconds(5));
    return modelPath;
}

inline void writeNpy(const std::string& path, const std::vector<float>& data, std::vector<uint64_t> shape) {
    int rc = sprl_write_npy_f32(path.c_str(), data.data(), shape.data(), (int)shape.size());
    if (rc != SPRL_OK) throw std::runtime_error("io error: failed to open a file.");      // utils/npy.hpp:633-636
}

template <typename NeuralNetwork, typename ImplNode, int NUM_ROWS, int NUM_COLS, int HISTORY_SIZE, int ACTION_SIZE>
void runWorker(std::string runName, std::string saveDir,
               INetwork<GridState<NUM_ROWS * NUM_COLS, HISTORY_SIZE>, ACTION_SIZE>* initialNetwork,
               ISymmetrizer<GridState<NUM_ROWS * NUM_COLS, HISTORY_SIZE>, ACTION_SIZE>* symmetrizer,
               int numIters,
               int initNumGamesPerWorker, int initUctTraversals, int initMaxBatchSize, int initMaxQueueSize,
               int numGamesPerWorker, int uctTraversals, int maxBatchSize, int maxQueueSize,
               float dirEps, float dirAlpha) {
    using State = GridState<NUM_ROWS * NUM_COLS, HISTORY_SIZE>;
    try {
        bool result = std::filesystem::create_directories(saveDir);
        std::cout << (result ? "Created directory: " : "Directory already exists: ") << saveDir << std::endl;
    } catch (std::exception& e) {
        std::cerr << "Error creating directory: " << e.what() << std::endl;
        return;
    }
    INetwork<State, ACTION_SIZE>* network;
    for (int iter = 0; iter < numIters; ++iter) {
        std::cout << "Starting iteration " << iter << "..." << std::endl;
        std::string modelPath = waitModelPath(iter - 1, runName);
        std::string savePath = saveDir + "/" + runName + "_iteration_" + std::to_string(iter);
        // The reference shadows maxBatchSize / maxQueueSize here and self-initialises them for
        // iter > 0 (GridWorker.hpp:120-121, undefined behaviour); the intended values are used.
        int numGames = (iter == 0) ? initNumGamesPerWorker : numGamesPerWorker;
        int numTraversals = (iter == 0) ? initUctTraversals : uctTraversals;
        int batch = (iter == 0) ? initMaxBatchSize : maxBatchSize;
        int queue = (iter == 0) ? initMaxQueueSize : maxQueueSize;

        NeuralNetwork neuralNetwork { modelPath };
        if (modelPath == "random") {
            std::cout << "Using initial network..." << std::endl;
            network = initialNetwork;
        } else {
            std::cout << "Using traced PyTorch network..." << std::endl;
            network = &neuralNetwork;
        }
        auto [states, distributions, outcomes] = runIteration<ImplNode, State, ACTION_SIZE>(
            network, numGames, numTraversals, batch, queue, dirEps, dirAlpha, InitQ::PARENT, symmetrizer, true);
        deviceOptions().firstGame += (uint64_t)numGames * deviceOptions().gameStride;

        const uint64_t n = outcomes.size();
        writeNpy(savePath + "_states.npy", states, { n, (uint64_t)(2 * HISTORY_SIZE + 1), (uint64_t)NUM_ROWS, (uint64_t)NUM_COLS });
        writeNpy(savePath + "_distributions.npy", distributions, { n, (uint64_t)ACTION_SIZE });
        writeNpy(savePath + "_outcomes.npy", outcomes, { n });
    }
}

}  // namespace SPRL
#endif
