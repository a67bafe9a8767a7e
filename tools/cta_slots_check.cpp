// Host replay of the tile schedule of krig_nll_cta.cu against its slot packing (csrc/cta_slots.h): for every NT, walk the block
// columns in the order the kernel does -- iteration J reads the tiles of the k-steps K < J of columns J and J + 1 (look-ahead),
// then stores the tiles (I, J), I > J -- and check that a store never lands on a slot whose tile is still to be read, that no two
// live tiles share a slot, and that the slot count is the one the launcher sizes shared memory with.
// Returns 0 when every NT in [lo, hi] passes, otherwise the first failing NT.
#include <vector>

#include "../surrogate-models_b200/csrc/cta_slots.h"

extern "C" int cta_slots_check(int lo, int hi, int* slots_out) {
    for (int NT = lo; NT <= hi; ++NT) {
        const int nslots = smb::cta_slots(NT);
        if (slots_out) slots_out[NT - lo] = nslots;
        std::vector<int> ownerI(nslots, -1), ownerK(nslots, -1);  // tile currently held by each slot
        // last iteration in which tile (I, K) is read: as an operand of column I (k-step K): K < I - 1 in the look-ahead of
        // iteration I - 1 (or, for the ex-owner's catch-up, at the top of iteration I), K = I - 1 at the top of iteration I;
        // as an A operand of column J (K < J <= I) never later than that.  So: dead after iteration I.
        for (int J = 0; J < NT; ++J) {
            // reads of iteration J: columns J and J + 1, k-steps K < J, rows I >= that column
            for (int JJ = J; JJ <= J + 1 && JJ < NT; ++JJ)
                for (int K = 0; K < (JJ == J ? J : J); ++K)
                    for (int I = JJ; I < NT; ++I) {
                        const int s = smb::cta_slot_of(NT, I, K);
                        if (s < 0 || s >= nslots || ownerI[s] != I || ownerK[s] != K) return NT;  // the tile must still be there
                    }
            // stores of iteration J: tiles (I, J), I > J
            for (int I = J + 1; I < NT; ++I) {
                const int s = smb::cta_slot_of(NT, I, J);
                if (s < 0 || s >= nslots) return NT;
                if (ownerI[s] >= 0 && ownerI[s] > J) return NT;  // the previous tenant (I', K') is read until iteration I'
                ownerI[s] = I;
                ownerK[s] = J;
            }
            // the owner of diagonal tile J defers its look-ahead: it reads column J + 1's k-steps K < J only now, after the stores
            if (J + 1 < NT)
                for (int K = 0; K < J; ++K)
                    for (int I = J + 1; I < NT; ++I) {
                        const int s = smb::cta_slot_of(NT, I, K);
                        if (ownerI[s] != I || ownerK[s] != K) return NT;
                    }
        }
    }
    return 0;
}
