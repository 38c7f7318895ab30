/*
 * sprl_b200.h -- C ABI of libsprl_b200.so, the B200-native self-play engine.
 *
 * The reference (willwin4sure/sprl) has no FFI: its hot path is a C++ template
 * library (cpp/src/{games,uct,networks,symmetry,selfplay}) driven by
 * cpp/src/OTHWorker.cpp.  This header is the boundary a maintainer binds
 * instead of instantiating those templates; each entry point names the
 * reference interface it replaces (paths relative to /root/reference/cpp/src).
 *
 * Conventions: plain pointers and sizes, no C++ or torch types.  Every function
 * returns 0 on success or a negative SPRL_E_* code; sprl_last_error() gives the
 * message of the last failure on the calling thread.  Pointers named h_* are
 * host memory, d_* device memory of the engine's GPU.  An engine handle is not
 * thread-safe; use one host thread per engine (= per GPU).
 */
#ifndef SPRL_B200_H
#define SPRL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPRL_OK 0
#define SPRL_E_INVALID (-1)    /* bad argument */
#define SPRL_E_CUDA (-2)       /* CUDA runtime error (sticky on the engine) */
#define SPRL_E_CAPACITY (-3)   /* a caller buffer or a device pool is too small */
#define SPRL_E_STATE (-4)      /* call not valid in the engine's current state */
#define SPRL_E_IO (-5)         /* file could not be written */
#define SPRL_E_NOGPU (-6)      /* no CUDA device: there is no CPU fallback */

/* games/{OthelloNode,ConnectFourNode,GoNode}.hpp */
#define SPRL_GAME_OTHELLO 0
#define SPRL_GAME_C4 1
#define SPRL_GAME_GO7 2
#define SPRL_GAME_GO9 3

/* networks/INetwork.hpp implementations */
#define SPRL_EVAL_UNIFORM 0    /* networks/RandomNetwork.hpp:21-49, evaluated on device */
#define SPRL_EVAL_HASHNET 1    /* deterministic test evaluator, evaluated on device */
#define SPRL_EVAL_EXTERNAL 2   /* networks/GridNetwork.hpp:62-145: caller runs the traced network on device buffers */

/* uct/UCTNode.hpp:24-28 */
#define SPRL_INITQ_ZERO 0
#define SPRL_INITQ_PARENT 1

const char* sprl_last_error(void);
int sprl_device_count(void);

typedef struct {
    int rows, cols, cells, actions, history, nsym, max_plies;
} sprl_game_info;

/* Compile-time constants of a game (games/OthelloNode.hpp:8-11 etc.). */
int sprl_game_info_get(int game, sprl_game_info* out);

/* ------------------------------------------------------------------ environment
 * Batched rules without a tree: replaces GameNode::getAddChild / getActionMask /
 * isTerminal / getWinner (games/GameNode.hpp:96-160) over many positions. */

/* One transition per position.  Inputs (host): cells [n, cells] int8 with the
 * reference's Piece encoding (-1 empty, 0, 1), player [n] (0/1), action [n].
 * Outputs (host): next cells, next player, terminal, winner (-1/0/1), mask
 * [n, actions] (0/1).  Othello and Connect Four only (Go needs the history path:
 * use sprl_env_rollout / sprl_env_perft). */
int sprl_env_step(int device, int game, int64_t n, const int8_t* h_cells, const int8_t* h_player,
                  const int32_t* h_action, int8_t* h_next_cells, int8_t* h_next_player,
                  int8_t* h_terminal, int8_t* h_winner, int8_t* h_mask);

/* Random playouts from the start position, one GPU thread per game; move k of
 * game g is legal[UniformInt(0, nlegal-1)] from the stream (seed, first_game+g).
 * h_game_steps [ngames] receives positions per game (start and terminal
 * included), h_final_winner [ngames] the winner.  When h_cells != NULL every
 * position is recorded in playing order (game-major), up to `cap` positions:
 * cells [cap, cells], player/terminal/winner [cap], mask [cap, actions],
 * action [cap] (-1 at terminal positions).  total_positions receives the count.
 * elapsed_ms (optional) receives the device time of the playout kernel alone. */
int sprl_env_rollout(int device, int game, uint64_t seed, uint64_t first_game, int64_t ngames,
                     int32_t* h_game_steps, int8_t* h_final_winner, int64_t cap,
                     int8_t* h_cells, int8_t* h_player, int8_t* h_terminal, int8_t* h_winner,
                     int8_t* h_mask, int32_t* h_action, int64_t* total_positions, float* elapsed_ms);

/* Leaf-count perft from the start position by breadth-first frontier expansion
 * on the device (a pass is a ply; a terminal node above the horizon counts 1). */
int sprl_env_perft(int device, int game, int depth, uint64_t* count, float* elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif
