// Host build of csrc/bingcd.cuh for tests/test_bingcd.py (the same code the GPU runs).
#include "../../go-eth-kzg_b200/csrc/bingcd.cuh"
extern "C" int bingcd_inverse(int nlimbs, const uint32_t *y, const uint32_t *m, uint32_t m_ninv31, uint32_t *out, int count) {
    int worst = 0;
    for (int i = 0; i < count; ++i) {
        int r = nlimbs == 12 ? kzg::BinGcd<12>::inverse(out + 12 * i, y + 12 * i, m, m_ninv31)
                             : kzg::BinGcd<8>::inverse(out + 8 * i, y + 8 * i, m, m_ninv31);
        if (r > worst) worst = r;
    }
    return worst;
}
