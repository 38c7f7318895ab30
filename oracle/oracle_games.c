/*
 * oracle_games.c -- TEST INFRASTRUCTURE (CPU oracle); see oracle.h.
 *
 * Rules of Othello, Connect Four and Go restated on int8 cell boards, with the
 * reference's conventions (who moves first, the pass action, what a terminal
 * node's mask looks like).  Go is restated from the RULES the reference
 * implements (suicide illegal, positional superko on the board only, two
 * passes or a depth cap end the game, Tromp-Taylor area score) with flood
 * fills and exact board comparison instead of its DSU + Zobrist machinery;
 * tests pin it to the verbatim reference through perft and rollout traces.
 */
#include "oracle_internal.h"

static const ogame_info k_info[4] = {
    /* rows cols cells actions history nsym max_plies komi */
    { 8, 8, 64, 65, 1, 8, 0, 0.0f },      /* games/OthelloNode.hpp:8-11 */
    { 6, 7, 42, 7, 1, 2, 0, 0.0f },       /* games/ConnectFourNode.hpp:8-13 */
    { 7, 7, 49, 50, 8, 8, 98, 9.0f },     /* games/GoNode.hpp:16-22 */
    { 9, 9, 81, 82, 8, 8, 162, 7.5f },    /* two-constant variant, games/GoDesc.md:127-128 */
};

const ogame_info* og_info(int game) { return (game >= 0 && game < 4) ? &k_info[game] : 0; }

int oracle_game_info(int game, ogame_info* out) {
    const ogame_info* gi = og_info(game);
    if (!gi) return -1;
    *out = *gi;
    return 0;
}

/* ------------------------------------------------------------------ Othello */
static const int k_dr[8] = { 1, 1, 0, -1, -1, -1, 0, 1 };   /* games/OthelloNode.cpp:199-200 */
static const int k_dc[8] = { 0, 1, 1, 1, 0, -1, -1, -1 };

static int oth_in(int r, int c) { return r >= 0 && r < 8 && c >= 0 && c < 8; }

/* games/OthelloNode.cpp:226-252: some direction holds >=1 opponent stones then ours */
static int oth_can_capture(const int8_t* b, int row, int col, int piece) {
    int opp = 1 - piece;
    for (int d = 0; d < 8; ++d) {
        int r = row + k_dr[d], c = col + k_dc[d], seen = 0;
        while (oth_in(r, c) && b[r * 8 + c] == opp) { r += k_dr[d]; c += k_dc[d]; seen = 1; }
        if (seen && oth_in(r, c) && b[r * 8 + c] == piece) return 1;
    }
    return 0;
}

/* games/OthelloNode.cpp:156-177 */
static void oth_mask(const int8_t* b, int player, float* mask) {
    int any = 0;
    for (int i = 0; i < 65; ++i) mask[i] = 0.0f;
    for (int r = 0; r < 8; ++r)
        for (int c = 0; c < 8; ++c) {
            if (b[r * 8 + c] != O_NONE) continue;
            if (oth_can_capture(b, r, c, player)) { mask[r * 8 + c] = 1.0f; any = 1; }
        }
    mask[64] = any ? 0.0f : 1.0f;
}

void og_othello_mask(const int8_t* b, int player, float* mask) { oth_mask(b, player, mask); }

/* games/OthelloNode.cpp:179-191 */
static int oth_terminal(const int8_t* b) {
    float m[65];
    oth_mask(b, O_ZERO, m);
    if (m[64] == 0.0f) return 0;
    oth_mask(b, O_ONE, m);
    return m[64] > 0.0f;
}

static void oth_start(onode* n) {       /* games/OthelloNode.cpp:18-32 */
    memset(n->cells, O_NONE, 64);
    n->cells[3 * 8 + 3] = O_ONE;
    n->cells[3 * 8 + 4] = O_ZERO;
    n->cells[4 * 8 + 3] = O_ZERO;
    n->cells[4 * 8 + 4] = O_ONE;
    n->player = O_ZERO;
    oth_mask(n->cells, n->player, n->mask);
}

static void oth_next(const onode* p, int action, onode* n) {   /* games/OthelloNode.cpp:34-87 */
    int piece = p->player, opp = 1 - piece;
    memcpy(n->cells, p->cells, 64);
    if (action != 64) {
        int row = action / 8, col = action % 8;
        n->cells[action] = (int8_t)piece;
        for (int d = 0; d < 8; ++d) {       /* flips, games/OthelloNode.cpp:193-224 */
            int r = row + k_dr[d], c = col + k_dc[d];
            while (oth_in(r, c) && n->cells[r * 8 + c] == opp) { r += k_dr[d]; c += k_dc[d]; }
            if (oth_in(r, c) && n->cells[r * 8 + c] == piece) {
                for (int rr = row + k_dr[d], cc = col + k_dc[d]; rr != r || cc != c; rr += k_dr[d], cc += k_dc[d])
                    n->cells[rr * 8 + cc] = (int8_t)piece;
            }
        }
    }
    n->player = (int8_t)opp;
    n->terminal = (int8_t)oth_terminal(n->cells);
    n->winner = O_NONE;
    if (n->terminal) {
        int c0 = 0, c1 = 0;
        for (int i = 0; i < 64; ++i) { c0 += n->cells[i] == O_ZERO; c1 += n->cells[i] == O_ONE; }
        if (c0 > c1) n->winner = O_ZERO;
        if (c1 > c0) n->winner = O_ONE;
    }
    oth_mask(n->cells, n->player, n->mask);
}

/* ------------------------------------------------------------- Connect Four */
static void c4_start(onode* n) {        /* games/ConnectFourNode.cpp:13-21 */
    memset(n->cells, O_NONE, 42);
    for (int i = 0; i < 7; ++i) n->mask[i] = 1.0f;
    n->player = O_ZERO;
}

static int c4_run(const int8_t* b, int row, int col, int dr, int dc, int piece) {
    int count = 1;
    for (int s = -1; s <= 1; s += 2) {
        int r = row + s * dr, c = col + s * dc;
        while (r >= 0 && r < 6 && c >= 0 && c < 7 && b[r * 7 + c] == piece) { ++count; r += s * dr; c += s * dc; }
    }
    return count;
}

static void c4_next(const onode* p, int action, onode* n) {    /* games/ConnectFourNode.cpp:23-78 */
    int piece = p->player;
    memcpy(n->cells, p->cells, 42);
    memcpy(n->mask, p->mask, 7 * sizeof(float));
    int col = action, row = 5;                    /* row 0 is the top */
    while (row >= 0 && n->cells[row * 7 + col] != O_NONE) --row;
    n->cells[row * 7 + col] = (int8_t)piece;
    if (row == 0) n->mask[col] = 0.0f;
    /* games/ConnectFourNode.cpp:135-217: four axes through the placed stone */
    int win = c4_run(n->cells, row, col, 0, 1, piece) >= 4 || c4_run(n->cells, row, col, 1, 0, piece) >= 4 ||
              c4_run(n->cells, row, col, 1, 1, piece) >= 4 || c4_run(n->cells, row, col, 1, -1, piece) >= 4;
    int filled = 1;
    for (int c = 0; c < 7; ++c) if (n->cells[c] == O_NONE) { filled = 0; break; }
    n->winner = win ? (int8_t)piece : O_NONE;
    n->terminal = (int8_t)(win || filled);
    if (n->terminal) for (int i = 0; i < 7; ++i) n->mask[i] = 0.0f;
    n->player = (int8_t)(1 - piece);
}

/* ----------------------------------------------------------------------- Go */
static int go_neighbors(int w, int coord, int out[4]) {    /* games/GoNode.hpp:92-104 */
    int r = coord / w, c = coord % w, k = 0;
    if (r > 0) out[k++] = coord - w;
    if (c > 0) out[k++] = coord - 1;
    if (r < w - 1) out[k++] = coord + w;
    if (c < w - 1) out[k++] = coord + 1;
    return k;
}

/* Flood the group of `coord`; returns its liberty count; group[] marks members. */
static int go_group(int w, const int8_t* b, int coord, uint8_t* group) {
    int n = w * w, piece = b[coord], libs = 0, head = 0, tail = 0;
    int queue[OG_MAXB];
    uint8_t seen[OG_MAXB];
    memset(seen, 0, (size_t)n);
    memset(group, 0, (size_t)n);
    seen[coord] = 1; group[coord] = 1; queue[tail++] = coord;
    while (head < tail) {
        int cur = queue[head++], nb[4];
        int k = go_neighbors(w, cur, nb);
        for (int i = 0; i < k; ++i) {
            int x = nb[i];
            if (seen[x]) continue;
            if (b[x] == piece) { seen[x] = 1; group[x] = 1; queue[tail++] = x; }
            else if (b[x] == O_NONE) { seen[x] = 1; ++libs; }
        }
    }
    return libs;
}

/* Place a stone and remove captured enemy groups (games/GoNode.cpp:96-176). */
static void go_place(int w, int8_t* b, int coord, int piece) {
    int nb[4], opp = 1 - piece;
    uint8_t group[OG_MAXB];
    b[coord] = (int8_t)piece;
    int k = go_neighbors(w, coord, nb);
    for (int i = 0; i < k; ++i) {
        if (b[nb[i]] != opp) continue;
        if (go_group(w, b, nb[i], group) == 0)
            for (int j = 0; j < w * w; ++j) if (group[j]) b[j] = O_NONE;
    }
}

/* games/GoNode.cpp:178-228: empty point, not suicide, and the resulting board
 * must differ from every earlier board on the path (positional superko on the
 * board only; passes add nothing; the empty start board can never recur). */
static int go_legal(int w, const onode* node, int coord, int piece) {
    int n = w * w;
    if (node->cells[coord] != O_NONE) return 0;
    int8_t b[OG_MAXB];
    uint8_t group[OG_MAXB];
    memcpy(b, node->cells, (size_t)n);
    go_place(w, b, coord, piece);
    if (go_group(w, b, coord, group) == 0) return 0;
    for (const onode* a = node; a; a = a->parent)
        if (memcmp(a->cells, b, (size_t)n) == 0) return 0;
    return 1;
}

static void go_score(int w, const int8_t* b, int terr[2]) {   /* games/GoNode.cpp:230-290 */
    int n = w * w;
    uint8_t seen[OG_MAXB];
    memset(seen, 0, (size_t)n);
    terr[0] = terr[1] = 0;
    for (int i = 0; i < n; ++i) {
        if (b[i] == O_ZERO) { terr[0]++; continue; }
        if (b[i] == O_ONE) { terr[1]++; continue; }
        if (seen[i]) continue;
        int queue[OG_MAXB], head = 0, tail = 0, count = 0, touch0 = 0, touch1 = 0;
        seen[i] = 1; queue[tail++] = i;
        while (head < tail) {
            int cur = queue[head++], nb[4];
            ++count;
            int k = go_neighbors(w, cur, nb);
            for (int j = 0; j < k; ++j) {
                int x = nb[j];
                if (b[x] == O_ZERO) touch0 = 1;
                else if (b[x] == O_ONE) touch1 = 1;
                else if (!seen[x]) { seen[x] = 1; queue[tail++] = x; }
            }
        }
        /* possibleTerritory[0] = !touch1, possibleTerritory[1] = !touch0 */
        if (!touch1 && touch0) terr[0] += count;
        if (!touch0 && touch1) terr[1] += count;
    }
}

static void go_start(const ogame_info* gi, onode* n) {      /* games/GoNode.cpp:303-317 */
    memset(n->cells, O_NONE, (size_t)gi->cells);
    for (int i = 0; i < gi->actions; ++i) n->mask[i] = 1.0f;
    n->player = O_ZERO;
    n->depth = 0;
}

static void go_next(const ogame_info* gi, const onode* p, int action, onode* n) {  /* games/GoNode.cpp:319-383 */
    int w = gi->cols, nc = gi->cells, piece = p->player;
    memcpy(n->cells, p->cells, (size_t)nc);
    if (action != nc) go_place(w, n->cells, action, piece);
    n->player = (int8_t)(1 - piece);
    n->depth = p->depth + 1;
    /* p->action is 0 at the root, so the root never counts as a pass */
    n->terminal = (int8_t)((p->action == nc && action == nc) || n->depth >= gi->max_plies);
    n->winner = O_NONE;
    for (int i = 0; i < gi->actions; ++i) n->mask[i] = 0.0f;
    if (!n->terminal) {
        for (int i = 0; i < nc; ++i) n->mask[i] = go_legal(w, n, i, n->player) ? 1.0f : 0.0f;
        n->mask[nc] = 1.0f;
    } else {
        int terr[2];
        go_score(w, n->cells, terr);
        float s0 = (float)terr[0], s1 = (float)terr[1];
        s1 += gi->komi;
        if ((double)s0 > (double)s1 + 0.1) n->winner = O_ZERO;
        else if ((double)s1 > (double)s0 + 0.1) n->winner = O_ONE;
    }
}

/* ------------------------------------------------------------- node plumbing */
static onode* node_alloc(void) {
    onode* n = (onode*)calloc(1, sizeof(onode));
    if (!n) abort();
    n->winner = O_NONE;
    return n;
}

onode* og_new_root(int game) {
    const ogame_info* gi = og_info(game);
    onode* n = node_alloc();
    switch (game) {
    case OG_OTHELLO: oth_start(n); break;
    case OG_C4: c4_start(n); break;
    default: go_start(gi, n); break;
    }
    return n;
}

/* games/GameNode.hpp:96-105.  For Go the new node must be linked to its parent
 * before the mask is computed (superko walks the path). */
onode* og_get_add_child(int game, onode* p, int action) {
    if (p->child[action]) return p->child[action];
    const ogame_info* gi = og_info(game);
    onode* n = node_alloc();
    n->parent = p;
    n->action = action;
    switch (game) {
    case OG_OTHELLO: oth_next(p, action, n); break;
    case OG_C4: c4_next(p, action, n); break;
    default: go_next(gi, p, action, n); break;
    }
    p->child[action] = n;
    return n;
}

void og_free_subtree(onode* n) {
    if (!n) return;
    for (int a = 0; a < OG_MAXA; ++a) og_free_subtree(n->child[a]);
    free(n);
}

void og_prune_children_except(onode* n, int keep, int nactions) {   /* games/GameNode.hpp:113-122 */
    for (int a = 0; a < nactions; ++a)
        if (a != keep && n->child[a]) { og_free_subtree(n->child[a]); n->child[a] = 0; }
}

void og_rewards(const onode* n, float out[2]) {     /* e.g. games/OthelloNode.cpp:94-100 */
    out[0] = out[1] = 0.0f;
    if (n->winner == O_ZERO) { out[0] = 1.0f; out[1] = -1.0f; }
    if (n->winner == O_ONE) { out[0] = -1.0f; out[1] = 1.0f; }
}

/* getGameStateImpl: Othello / C4 a single board; Go walks parent pointers for up
 * to 8 boards, through the pre-root history too (games/GoNode.cpp:385-398). */
int og_game_state(int game, const onode* n, int8_t hist[OG_MAXH][OG_MAXB]) {
    const ogame_info* gi = og_info(game);
    int t = 0;
    const onode* cur = n;
    while (t < gi->history && cur) {
        memcpy(hist[t], cur->cells, (size_t)gi->cells);
        cur = cur->parent;
        ++t;
    }
    return t;
}

/* ------------------------------------------------------------ env-only API */
static uint64_t perft_rec(int game, onode* n, int depth, int nactions) {
    if (depth == 0 || n->terminal) return 1;
    uint64_t total = 0;
    for (int a = 0; a < nactions; ++a) {
        if (n->mask[a] == 0.0f) continue;
        onode* c = og_get_add_child(game, n, a);
        total += perft_rec(game, c, depth - 1, nactions);
        og_free_subtree(c);
        n->child[a] = 0;
    }
    return total;
}

int oracle_perft(int game, int depth, uint64_t* count) {
    const ogame_info* gi = og_info(game);
    if (!gi || depth < 0) return -1;
    onode* root = og_new_root(game);
    *count = perft_rec(game, root, depth, gi->actions);
    og_free_subtree(root);
    return 0;
}

int64_t oracle_rollout(int game, uint64_t seed, uint64_t first_game, int ngames, int64_t cap,
                       int32_t* game_steps, int8_t* cells, int8_t* player, int8_t* terminal,
                       int8_t* winner, int8_t* mask, int32_t* action) {
    const ogame_info* gi = og_info(game);
    if (!gi) return -1;
    int64_t pos = 0;
    for (int g = 0; g < ngames; ++g) {
        orng_t rng = { seed, first_game + (uint64_t)g, 0 };
        onode* root = og_new_root(game);
        onode* cur = root;
        int steps = 0;
        for (;;) {
            if (pos >= cap) { og_free_subtree(root); return -1; }
            int legal[OG_MAXA], nl = 0;
            memcpy(cells + pos * gi->cells, cur->cells, (size_t)gi->cells);
            player[pos] = cur->player;
            terminal[pos] = cur->terminal;
            winner[pos] = cur->winner;
            for (int a = 0; a < gi->actions; ++a) {
                int ok = cur->mask[a] != 0.0f;
                mask[pos * gi->actions + a] = (int8_t)ok;
                if (ok) legal[nl++] = a;
            }
            ++steps;
            if (cur->terminal) { action[pos++] = -1; break; }
            int pick = legal[orng_uniform_int(&rng, 0, nl - 1)];
            action[pos++] = pick;
            cur = og_get_add_child(game, cur, pick);
        }
        game_steps[g] = steps;
        og_free_subtree(root);
    }
    return pos;
}

int oracle_replay(int game, const int32_t* actions, int n, int8_t* cells, int8_t* player,
                  int8_t* terminal, int8_t* winner, int8_t* mask, float* rewards2) {
    const ogame_info* gi = og_info(game);
    if (!gi) return -1;
    onode* root = og_new_root(game);
    onode* cur = root;
    for (int i = 0; i < n; ++i) {
        if (cur->terminal || actions[i] < 0 || actions[i] >= gi->actions || cur->mask[actions[i]] == 0.0f) {
            og_free_subtree(root);
            return -2 - i;          /* illegal move at index i */
        }
        cur = og_get_add_child(game, cur, actions[i]);
    }
    if (cells) memcpy(cells, cur->cells, (size_t)gi->cells);
    if (player) *player = cur->player;
    if (terminal) *terminal = cur->terminal;
    if (winner) *winner = cur->winner;
    if (mask) for (int a = 0; a < gi->actions; ++a) mask[a] = (int8_t)(cur->mask[a] != 0.0f);
    if (rewards2) og_rewards(cur, rewards2);
    og_free_subtree(root);
    return 0;
}
