#!/usr/bin/env python
"""Generate the committed golden fixtures (run HERE, where /root/reference exists; never on the GPU box).

Inputs come from the reference's own bundled data (dataset/kinect/*.png, depth.yuv) pushed through the
pre-codec steps of simpleTest()/pointcloudTest() (main.cpp:507-539, :255-285) with the reference build
(oracle/_ref/libdmc_ref.so) and cv2's JPEG codec.  The decoded 8-bit disparity images are stored as PNG
(lossless) so that the GPU box does not depend on a JPEG decoder version; expected outputs are stored as
CRC-32 of the raw row-major bytes, all produced by the UNMODIFIED reference sources.

The first block re-derives the known answers recorded in SURVEY.md section 8c and aborts on any drift.
"""
import json
import os
import sys
import zlib

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.oracle_py import Reference, FILL_DISPARITY, SEPARABLE_KERNEL, FULL_KERNEL  # noqa: E402

REFDIR = "/root/reference/PostFilterSetForDepthCoding/"
FOCUS, BASELINE, AMP = 75.0, 575.0, 2.6


def crc(a):
    return "%08x" % (zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF)


SURVEY_KAT = {   # SURVEY.md 8c "Probe known answers"
    "meeting_small_1_1": dict(disp="e81c9bd4", fill1="c5e6baf9", dec50="1dabb0f4", pfs="0874aa69", fill2="b3195451",
                              dec80="4fe53ca8", d32f="c09ece5a", d2d="ac6f78a1", xyz="20266593", d16u="2f17dc46",
                              disp32f="739a77b8", brf="14887a4d"),
    "desk_1_1": dict(disp="871b406e", fill1="6c83c5de", dec50="a35a1bae", pfs="3906d9ff", fill2="1f383d43",
                     dec80="7ed99945", d32f="0b7d7035", d2d="414d203c", xyz="1f44a729", d16u="6c75d013",
                     disp32f="521e47f0", brf="5c5df4f4"),
    "x264": dict(y="c449a178", pfs="22036da9", d32f="4ce35acf", brf="a7699b51", sep="ad276931"),
}


def chain_goldens(R, img):
    """Expected CRCs for one decoded 8-bit disparity image."""
    g = {}
    g["pfs_2_1_3_5_10"] = crc(R.post_filter_set(img, 2, 1, 3, 5, 10))
    g["pfs_1_0_1_3_10"] = crc(R.post_filter_set(img, 1, 0, 1, 3, 10))
    g["pfs_2_1_3_5_10_sep"] = crc(R.post_filter_set(img, 2, 1, 3, 5, 10, SEPARABLE_KERNEL))
    d32 = R.filter_disp8u_depth32f(img, FOCUS, BASELINE, AMP, 1, 0, 1, 3, 65.0)
    g["depth32f_1_0_1_3_65"] = crc(d32)
    g["depth16u_1_0_1_3_65"] = crc(R.filter_disp8u_depth16u(img, FOCUS, BASELINE, AMP, 1, 0, 1, 3, 65.0))
    g["disp32f_1_0_1_3_10"] = crc(R.filter_disp8u_disp32f(img, 1, 0, 1, 3, 10.0))
    g["depth32f2disp8u"] = crc(R.depth32f2disp8u(d32, FOCUS * BASELINE, AMP, 0.0))
    g["reproject_xyz_510"] = crc(R.reproject_xyz(d32, 510.0))
    g["brf_13_1_1_1"] = crc(R.brf(img, 13, 13, 1, 1, 1))
    g["brf_7_1_2_05"] = crc(R.brf(img, 7, 7, 1, 2, 0.5))
    for r in range(1, 8):
        g["bwrf8u_r%d_th10" % r] = crc(R.bwrf(img, 2 * r + 1, 2 * r + 1, 10))
    g["bwrf8u_sep_11_th10"] = crc(R.bwrf(img, 11, 11, 10, SEPARABLE_KERNEL))
    f = img.astype(np.float32) * 16.0
    g["bwrf32f_r3_th65"] = crc(R.bwrf(f, 7, 7, 65.0))
    g["bwrf32f_r5_th65"] = crc(R.bwrf(f, 11, 11, 65.0))          # exercises the rpad=-1 padding quirk (cols%4==0)
    u16 = (img.astype(np.uint16) * 16)
    g["bwrf16u_r2_th160"] = crc(R.bwrf(u16, 5, 5, 160.0))
    for r in (1, 3):
        g["minmax_r%d" % r] = crc(R.blur_remove_minmax(img, r))
    for k in (3, 5):
        g["median_k%d" % k] = crc(R.median_blur(img, k))
    for gr in (1, 2):
        g["gauss_gr%d" % gr] = crc(R.small_gaussian(img, 2 * gr + 1, gr + 0.5))
    g["disp8u2depth32f"] = crc(R.disp8u2depth32f(img, FOCUS * BASELINE, AMP, 0.0))
    return g


def main():
    R = Reference()
    gold = {"_doc": "CRC-32 of raw row-major output bytes from oracle/_ref/libdmc_ref.so (unmodified reference); see make_golden.py",
            "inputs": {}}
    for name in ("meeting_small_1_1", "desk_1_1"):
        kat = SURVEY_KAT[name]
        d16 = cv2.imread(REFDIR + "dataset/kinect/%s_depth.png" % name, cv2.IMREAD_UNCHANGED)
        disp = R.depth16u2disp8u(d16, FOCUS * BASELINE, AMP, 0.0)                       # main.cpp:511
        assert crc(disp) == kat["disp"]
        f1 = R.fill_occlusion(disp, 0, FILL_DISPARITY)                                  # main.cpp:512
        assert crc(f1) == kat["fill1"]
        ok, buf = cv2.imencode(".jpg", f1, [cv2.IMWRITE_JPEG_QUALITY, 50])             # main.cpp:516-519
        dec50 = cv2.imdecode(buf, 0)
        assert crc(dec50) == kat["dec50"], "cv2 JPEG codec drift"
        t = np.ascontiguousarray(f1.T); t = R.fill_occlusion(t, 0, FILL_DISPARITY)      # main.cpp:257-260
        f2 = np.ascontiguousarray(t.T)
        assert crc(f2) == kat["fill2"]
        ok, buf = cv2.imencode(".jpg", f2, [cv2.IMWRITE_JPEG_QUALITY, 80])
        dec80 = cv2.imdecode(buf, 0)
        assert crc(dec80) == kat["dec80"]
        assert crc(R.post_filter_set(dec50, 2, 1, 3, 5, 10)) == kat["pfs"]
        d32 = R.filter_disp8u_depth32f(dec80, FOCUS, BASELINE, AMP, 1, 0, 1, 3, 65.0)
        assert crc(d32) == kat["d32f"]
        assert crc(R.depth32f2disp8u(d32, FOCUS * BASELINE, AMP, 0.0)) == kat["d2d"]
        assert crc(R.reproject_xyz(d32, 510.0)) == kat["xyz"]
        assert crc(R.filter_disp8u_depth16u(dec80, FOCUS, BASELINE, AMP, 1, 0, 1, 3, 65.0)) == kat["d16u"]
        assert crc(R.filter_disp8u_disp32f(dec80, 1, 0, 1, 3, 10.0)) == kat["disp32f"]
        assert crc(R.brf(dec80, 13, 13, 1, 1, 1)) == kat["brf"]
        short = "meeting" if name.startswith("meeting") else "desk"
        for q, im in ((50, dec50), (80, dec80)):
            fn = "kinect_%s_q%d.png" % (short, q)
            cv2.imwrite(os.path.join(HERE, fn), im, [cv2.IMWRITE_PNG_COMPRESSION, 9])
            assert np.array_equal(cv2.imread(os.path.join(HERE, fn), cv2.IMREAD_UNCHANGED), im)
            gold["inputs"][fn] = {"crc": crc(im), "golden": chain_goldens(R, im)}
        if short == "meeting":   # pre-codec steps on a 320x240 crop of the raw 16-bit depth (zeros = invalid)
            crop = np.ascontiguousarray(d16[120:360, 160:480])
            fn = "kinect_meeting_depth16_crop.png"
            cv2.imwrite(os.path.join(HERE, fn), crop, [cv2.IMWRITE_PNG_COMPRESSION, 9])
            dc = R.depth16u2disp8u(crop, FOCUS * BASELINE, AMP, 0.0)
            fc = R.fill_occlusion(dc, 0, FILL_DISPARITY)
            tc = np.ascontiguousarray(fc.T); tc = R.fill_occlusion(tc, 0, FILL_DISPARITY)
            gold["inputs"][fn] = {"crc": crc(crop), "golden": {
                "depth16u2disp8u": crc(dc), "fill_disparity_1pass": crc(fc), "fill_disparity_2pass": crc(np.ascontiguousarray(tc.T)),
                # (FILL_DEPTH with invalid == 0 == edge value makes the reference read s[cols]: undefined, not a golden)
                "reproject_xyz_16u_510": crc(R.reproject_xyz(crop, 510.0))}}
    y = np.fromfile(REFDIR + "depth.yuv", np.uint8)[:640 * 480].reshape(480, 640).copy()   # x264-decoded disparity (x264FFMPEGDemo.cpp:22-35)
    kat = SURVEY_KAT["x264"]
    assert crc(y) == kat["y"]
    assert crc(R.post_filter_set(y, 2, 1, 3, 5, 10)) == kat["pfs"]
    assert crc(R.filter_disp8u_depth32f(y, FOCUS, BASELINE, AMP, 1, 0, 1, 3, 65.0)) == kat["d32f"]
    assert crc(R.brf(y, 13, 13, 1, 1, 1)) == kat["brf"]
    assert crc(R.bwrf(y, 11, 11, 10, SEPARABLE_KERNEL)) == kat["sep"]
    fn = "x264_depth_y.png"
    cv2.imwrite(os.path.join(HERE, fn), y, [cv2.IMWRITE_PNG_COMPRESSION, 9])
    gold["inputs"][fn] = {"crc": crc(y), "golden": chain_goldens(R, y)}
    gold["survey_8c_known_answers_reproduced"] = True
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("wrote", len(gold["inputs"]), "inputs")


if __name__ == "__main__":
    main()
