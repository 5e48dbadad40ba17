"""Device encoder (nbl_encoder_create / nbl_encode_batch) against the reference's own transmitted symbols, and the
random-message mode of the device Monte-Carlo path."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, load_product_code

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", [n for n in golden_names() if "allzero" not in n and "qam" not in n])
def test_encoder_reproduces_reference_codewords(name, code_files):
    """TX_MSG_SYM -> TX_CODE_SYM exactly as CNBLDPC::Encode produced them in the reference run (random messages)"""
    from nbldpc_b200 import capi, gfcode
    z, meta = load_golden(name)
    code = load_product_code(code_files, meta["code"])
    enc = capi.Encoder(code)
    assert enc.K == code.N - code.M
    assert enc.n_swaps == 0                # with swaps the golden message would be the post-swap one (NBLDPC.cpp:599-602)
    cw = enc.encode(z["tx_msg_sym"])
    assert np.array_equal(cw, z["tx_code_sym"].astype(np.int32))
    assert not gfcode.syndromes(code, cw).any()


def test_encoder_random_messages_are_codewords_and_systematic(code_files):
    from nbldpc_b200 import capi, gfcode
    rng = np.random.default_rng(5)
    for name in ("divsalar.UNBLDPC.128.64.GF.16", "BDS.576.288.GF.64", "divsalar.CNBLDPC.512.256.GF.256", "H_bit_GF256"):
        code = load_product_code(code_files, name)
        enc = capi.Encoder(code)
        msg = rng.integers(0, code.q, size=(257, enc.K)).astype(np.int32)
        cw = enc.encode(msg)
        assert not gfcode.syndromes(code, cw).any(), name
        if enc.n_swaps == 0:
            assert np.array_equal(cw[:, :enc.K], msg), name
        # linearity over GF(q): the sum of two code words is the code word of the summed messages
        assert np.array_equal(enc.encode(msg[:100] ^ msg[100:200]), cw[:100] ^ cw[100:200])


def test_monte_carlo_with_device_encoded_random_messages(code_files):
    """random_msg = 1: counters are independent of the batch split and statistically equal to the all-zero run"""
    from nbldpc_b200 import capi
    code = load_product_code(code_files, "divsalar.UNBLDPC.512.256.GF.256")
    kw = dict(method=4, max_iter=10, p0=2, p1=2, factor=1.0, offset=0.5, precision=capi.PREC_F32)
    a = capi.Decoder(code, max_batch=512, **kw).simulate(2.0, 2048, seed=3, random_msg=1)
    b = capi.Decoder(code, max_batch=300, **kw)
    cb = b.simulate(2.0, 1000, seed=3, first_frame=0, random_msg=1)
    cb = b.simulate(2.0, 1048, seed=3, first_frame=1000, random_msg=1, counters=cb)
    assert np.array_equal(a, cb)
    assert a[0] == 2048 and 0 < a[1] < 2048
    zero = capi.Decoder(code, max_batch=512, **kw).simulate(2.0, 2048, seed=3)
    # same noise, different code words: frame error counts within 5 sigma of each other
    pa, pz = a[1] / 2048.0, zero[1] / 2048.0
    sd = np.sqrt(max(pa, pz) * (1 - min(pa, pz)) / 2048.0 * 2)
    assert abs(pa - pz) < 5 * sd + 1e-9, (pa, pz)
