"""NBLDPC.Profile.txt writer for driving the compiled reference (test / bench infrastructure).

The reference reads 34 values positionally, each preceded by a fixed number of throw-away label tokens
(Simulation.cpp:57-108); only the label TOKEN COUNT matters.  The file itself is not shipped with the reference.
"""

PROFILE_KEYS = [  # (label, key) in file order -- Simulation.cpp:57-108; only the label TOKEN COUNT matters
    ("GFq:", "GFq"), ("NonBinary File:", "code_file"), ("Puncture VarDegree:", "punct"), ("Decode Method:", "method"),
    ("Max Iter:", "max_iter"), ("parallel:", "parallel"), ("crcLen:", "crc"), ("crc_correctLen:", "crc_correct"),
    ("OSD_order:", "osd_order"), ("OSD_factor:", "osd_factor"), ("OSD_flag:", "osd_flag"),
    ("EMS Nm:", "ems_nm"), ("EMS Nc:", "ems_nc"), ("EMS Factor:", "ems_factor"), ("EMS Offset:", "ems_offset"),
    ("TEMS Nr:", "tems_nr"), ("TEMS Nc:", "tems_nc"), ("TEMS Factor:", "tems_factor"), ("TEMS Offset:", "tems_offset"),
    ("BSTEMS Nm:", "bs_nm"), ("BSTEMS Nc:", "bs_nc"), ("BSTEMS Factor:", "bs_factor"), ("BSTEMS Offset:", "bs_offset"),
    ("SNR Begin:", "snr"), ("SNR Step:", "snr_step"), ("SNR Stop:", "snr_stop"), ("nQAM:", "nqam"),
    ("Constellation:", "constellation"), ("Random Msg:", "random_msg"), ("Min Err Frame:", "min_err"),
    ("Min UErr Frame:", "min_uerr"), ("Min Sim Cycle:", "min_cycle"), ("Random Seed:", "seed"),
    ("Show Frame Step:", "show_step")]

DEFAULTS = dict(punct=0, max_iter=10, parallel=1, crc=8, crc_correct=0, osd_order=-1, osd_factor=1, osd_flag=0,
                ems_nm=8, ems_nc=3, ems_factor=1, ems_offset=0.5, tems_nr=2, tems_nc=2, tems_factor=1, tems_offset=0.5,
                bs_nm=4, bs_nc=3, bs_factor=1, bs_offset=0, snr_step=1, nqam=2, constellation="BPSK.txt",
                random_msg=1, min_err=30, min_uerr=-1, min_cycle=300, seed=1, show_step=20)


def profile_text(cfg):
    d = dict(DEFAULTS); d.update(cfg); d["snr_stop"] = d["snr"]
    return "".join("%s %s\n" % (lab, d[key]) for lab, key in PROFILE_KEYS)
