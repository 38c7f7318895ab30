// hashnet.hpp -- TEST INFRASTRUCTURE.  Deterministic evaluator plugged into the
// verbatim reference as an INetwork (networks/INetwork.hpp:16-33), so that the
// reference's own UCTTree::evaluateAndBackpropLeaves (uct/UCTTree.hpp:124-184)
// drives it, symmetry quirk Q3 included.  The mask -> sum -> normalise tail is
// written with the reference's GameActionDist operators, as in
// networks/GridNetwork.hpp:117-139.
#ifndef SPRL_REF_HASHNET_HPP
#define SPRL_REF_HASHNET_HPP

#include "games/GridState.hpp"
#include "networks/INetwork.hpp"

#include "../oracle_rng.h"

namespace SPRLREF {

template <int BOARD_SIZE, int HISTORY_SIZE, int ACTION_SIZE>
class HashNet : public SPRL::INetwork<SPRL::GridState<BOARD_SIZE, HISTORY_SIZE>, ACTION_SIZE> {
public:
    using State = SPRL::GridState<BOARD_SIZE, HISTORY_SIZE>;
    using ActionDist = SPRL::GameActionDist<ACTION_SIZE>;

    explicit HashNet(uint64_t salt = 0) : m_salt(salt) {}

    std::vector<std::pair<ActionDist, SPRL::Value>> evaluate(
        const std::vector<State>& states, const std::vector<ActionDist>& masks) override {
        int n = (int)states.size();
        m_numEvals += n;
        std::vector<std::pair<ActionDist, SPRL::Value>> out;
        out.reserve(n);
        for (int b = 0; b < n; ++b) {
            const State& s = states[b];
            SPRL::Piece own = SPRL::pieceFromPlayer(s.getPlayer());
            SPRL::Piece opp = SPRL::otherPiece(own);
            uint64_t words[4 * HISTORY_SIZE] = {};
            for (int t = 0; t < s.size(); ++t) {
                for (int i = 0; i < BOARD_SIZE; ++i) {
                    SPRL::Piece p = s.getHistory()[t][i];
                    if (p == own) words[4 * t + (i >> 6)] |= 1ULL << (i & 63);
                    if (p == opp) words[4 * t + 2 + (i >> 6)] |= 1ULL << (i & 63);
                }
            }
            uint64_t h = ohashnet_salt(ohashnet_state_hash(words, s.size(), (int)s.getPlayer()), m_salt);
            ActionDist policy;
            for (int i = 0; i < ACTION_SIZE; ++i) policy[i] = ohashnet_prior_raw(h, i);
            int numLegal = 0;
            for (int i = 0; i < ACTION_SIZE; ++i) {
                if (masks[b][i] == 0.0f) policy[i] = 0.0f; else ++numLegal;
            }
            float sum = policy.sum();
            if (sum == 0.0f) {
                float uniform = 1.0f / numLegal;
                for (int i = 0; i < ACTION_SIZE; ++i) policy[i] = (masks[b][i] == 0.0f) ? 0.0f : uniform;
            } else {
                policy = policy / sum;
            }
            out.emplace_back(policy, ohashnet_value(h));
        }
        return out;
    }
    int getNumEvals() override { return m_numEvals; }

private:
    int m_numEvals { 0 };
    uint64_t m_salt;
};

}  // namespace SPRLREF
#endif
