// Live-range packing of the strictly-lower 8 x 8 tiles of a Cholesky factor held in shared memory (krig_nll_cta.cu).
// Tile (I, K), K < I, is created when block column K is solved and last read while block column I is assembled, so the band
// a = I - K - 1 of the tile grid never has more than a + 2 tiles alive: it gets a ring of min(a + 2, NT - 1 - a) slots and tile
// (I, K) sits in slot K mod ring size of its band.  Plain C++ so the host test (tests/test_host_logic.py, tools/cta_slots_check.cpp)
// can replay the kernel's schedule against it.
#pragma once
#ifdef __CUDACC__
#define SMB_HD __host__ __device__
#else
#define SMB_HD
#endif

namespace smb {

SMB_HD inline int cta_band_slots(int NT, int band) {
    const int x = band + 2, y = NT - 1 - band;
    return x < y ? x : y;
}
SMB_HD inline int cta_slots(int NT) {
    int s = 0;
    for (int b = 0; b + 1 < NT; ++b) s += cta_band_slots(NT, b);
    return s > 0 ? s : 1;
}
// slot index of tile (I, K), K < I
SMB_HD inline int cta_slot_of(int NT, int I, int K) {
    const int band = I - K - 1;
    int base = 0;
    for (int b = 0; b < band; ++b) base += cta_band_slots(NT, b);
    return base + K % cta_band_slots(NT, band);
}

}  // namespace smb
