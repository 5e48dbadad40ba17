"""fp32 fast mode against EVERY FER anchor the reference itself produced (the compiled reference, seed 1,
one CPU core): the device Monte-Carlo's frame error rate must lie inside the exact (Clopper-Pearson) 95 % confidence interval
computed from the reference's own error / frame counts.  north_star: "in the fp32 fast mode ... BER/FER must fall inside the
reference's 95% confidence interval at every Eb/N0 point".

The device run uses >= 2e5 frames per point (>= 2e6 where the FER is below 1e-3), so its own sampling error is small against
the width of the reference's interval (the reference stopped after 2 001 ... 54 920 frames)."""
import numpy as np
import pytest
from scipy.stats import beta

from conftest import load_product_code

pytestmark = pytest.mark.gpu


def clopper_pearson(k, n, conf=0.95):
    a = (1.0 - conf) / 2.0
    lo = 0.0 if k == 0 else float(beta.ppf(a, k, n - k + 1))
    hi = 1.0 if k == n else float(beta.ppf(1.0 - a, k + 1, n - k))
    return lo, hi


# (label, code, method, p0, p1, offset, Eb/N0 dB, random message?, reference frame errors, reference frames, device frames)
ANCHORS = [
    ("cfg1 QSPA GF(16) 1 dB", "divsalar.UNBLDPC.128.64.GF.16", 1, 0, 0, 0.0, 1.0, 0, 1147, 2001, 400000),
    ("cfg1 QSPA GF(16) 2 dB", "divsalar.UNBLDPC.128.64.GF.16", 1, 0, 0, 0.0, 2.0, 0, 343, 2001, 400000),
    ("cfg1 QSPA GF(16) 3 dB", "divsalar.UNBLDPC.128.64.GF.16", 1, 0, 0, 0.0, 3.0, 0, 51, 3380, 800000),
    ("cfg1 QSPA GF(16) 4 dB", "divsalar.UNBLDPC.128.64.GF.16", 1, 0, 0, 0.0, 4.0, 0, 21, 54920, 4000000),
    ("cfg2 EMS GF(64) 8/3/0.5 2 dB", "BDS.576.288.GF.64", 2, 8, 3, 0.5, 2.0, 0, 65, 2001, 300000),
    ("cfg3 T-EMS GF(256) 2/2/0 2.5 dB", "divsalar.UNBLDPC.512.256.GF.256", 4, 2, 2, 0.0, 2.5, 0, 31, 1173, 200000),
    ("cfg3 T-EMS GF(256) 2/2/0.5 2.5 dB", "divsalar.UNBLDPC.512.256.GF.256", 4, 2, 2, 0.5, 2.5, 0, 16, 1120, 200000),
    ("cfg3 T-EMS GF(256) 2/2/1.0 2.5 dB", "divsalar.UNBLDPC.512.256.GF.256", 4, 2, 2, 1.0, 2.5, 0, 14, 1120, 200000),
    ("cfg5 T-EMS H_bit_GF256 3 dB all-zero", "H_bit_GF256", 4, 2, 2, 0.5, 3.0, 0, 71, 2001, 400000),
    ("cfg5 T-EMS H_bit_GF256 3 dB random", "H_bit_GF256", 4, 2, 2, 0.5, 3.0, 1, 54, 2001, 400000),
]


@pytest.mark.parametrize("anchor", ANCHORS, ids=[a[0] for a in ANCHORS])
def test_fp32_fer_inside_the_reference_confidence_interval(anchor, code_files):
    from nbldpc_b200 import capi
    label, code_name, method, p0, p1, offset, snr, rnd, k, n, frames = anchor
    code = load_product_code(code_files, code_name)
    dec = capi.Decoder(code, method, 10, p0, p1, 1.0, offset, precision=capi.PREC_F32, early_stop=1, max_batch=50000)
    assert dec.kernel() != "group", "the anchor must run on the configuration's fast kernel"
    c = dec.simulate(snr, frames, seed=1, random_msg=rnd, crc_len=8 if rnd else 0)
    fer = int(c[1]) / int(c[0])
    lo, hi = clopper_pearson(k, n)
    print("%s: device FER %.4e (%d / %d, kernel %s); reference %d / %d = %.4e, 95 %% CI [%.4e, %.4e]" %
          (label, fer, int(c[1]), int(c[0]), dec.kernel(), k, n, k / n, lo, hi))
    assert int(c[0]) == frames
    assert lo <= fer <= hi, (label, fer, lo, hi)


def test_confidence_interval_helper():
    lo, hi = clopper_pearson(54, 2001)
    assert 0.0203 < lo < 0.0205 and 0.0350 < hi < 0.0352        # exact binomial interval of 54 / 2001
    assert clopper_pearson(0, 100)[0] == 0.0 and clopper_pearson(100, 100)[1] == 1.0
