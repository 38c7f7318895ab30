/* oracle_internal.h -- TEST INFRASTRUCTURE (CPU oracle); see oracle.h. */
#ifndef SPRL_ORACLE_INTERNAL_H
#define SPRL_ORACLE_INTERNAL_H

#include "oracle.h"
#include "oracle_rng.h"

#include <stdlib.h>
#include <string.h>

/* Piece / Player encoding of games/GameNode.hpp:17-21, games/GridState.hpp:18-22 */
#define O_NONE (-1)
#define O_ZERO 0
#define O_ONE 1

/* One node = the reference's GameNode (games/GameNode.hpp:50-200) fused with its
 * UCTNode (uct/UCTNode.hpp:38-391); the reference creates them 1:1
 * (uct/UCTNode.hpp:258-284). */
typedef struct onode {
    struct onode* parent;
    struct onode* child[OG_MAXA];
    int action;                 /* m_action: action into this node, 0 at the root */
    int8_t player, winner, terminal;
    float mask[OG_MAXA];        /* m_actionMask */
    int8_t cells[OG_MAXB];      /* m_board */
    int depth;                  /* Go: m_depth */
    /* UCT part */
    int expanded, evaluated;    /* m_isExpanded, m_isNetworkEvaluated */
    float net_policy[OG_MAXA];  /* m_networkPolicy */
    float net_value;            /* m_networkValue */
    float P[OG_MAXA], W[OG_MAXA], N[OG_MAXA];   /* m_edgeStatistics */
    float* own_N;               /* &m_parentEdgeStatistics->m_numVisits[m_action] */
    float* own_W;
} onode;

const ogame_info* og_info(int game);
onode* og_new_root(int game);
onode* og_get_add_child(int game, onode* n, int action);   /* GameNode::getAddChild */
void og_free_subtree(onode* n);
void og_prune_children_except(onode* n, int keep, int nactions);
void og_rewards(const onode* n, float out[2]);
/* GameNode::getGameState: hist[t][cell], returns valid length */
int og_game_state(int game, const onode* n, int8_t hist[OG_MAXH][OG_MAXB]);

/* OthelloNode::actionMask(board, player), games/OthelloNode.cpp:156-177 (65 entries) */
void og_othello_mask(const int8_t* b, int player, float* mask);

void osym_cells(int game, int sym, const int8_t* in, int8_t* out);
void osym_dist(int game, int sym, const float* in, float* out);
int osym_inverse(int game, int sym);

#endif
