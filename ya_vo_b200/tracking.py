"""Python mirror of the OpenCV call the reference's steady-state loop makes after feature extraction
(LoopHandler::trackLastFrame, src/LoopHandler.cc:372-375), over the C ABI (yavo_klt_track).

calcOpticalFlowPyrLK keeps cv2's signature, defaults and return shapes, so a test written against
cv2.calcOpticalFlowPyrLK runs against it unchanged:

    nextPts, status, err = calcOpticalFlowPyrLK(prevImg, nextImg, prevPts, None, winSize=(11, 11), maxLevel=3,
                                                criteria=(COUNT + EPS, 30, 0.01), flags=0, minEigThreshold=0.001)

Points are OpenCV's (x = column, y = row).  The pyramid (cv2.pyrDown chain) and the per-point iterations run in the
CUDA library; results are bit-identical to cv2 4.13 (tests/golden/klt_golden.npz).
"""
import numpy as np

from . import frontend

TERM_CRITERIA_COUNT, TERM_CRITERIA_EPS = 1, 2
OPTFLOW_USE_INITIAL_FLOW, OPTFLOW_LK_GET_MIN_EIGENVALS = 4, 8


def calcOpticalFlowPyrLK(prevImg, nextImg, prevPts, nextPts=None, winSize=(21, 21), maxLevel=3,
                         criteria=(TERM_CRITERIA_COUNT + TERM_CRITERIA_EPS, 30, 0.01), flags=0, minEigThreshold=1e-4,
                         ctx=None):
    prev = prevImg.rawImage if isinstance(prevImg, frontend.Image) else np.asarray(prevImg)
    nxt = nextImg.rawImage if isinstance(nextImg, frontend.Image) else np.asarray(nextImg)
    if prev.dtype != np.uint8 or prev.ndim != 2 or prev.shape != nxt.shape or nxt.dtype != np.uint8:
        raise ValueError("calcOpticalFlowPyrLK expects two 8-bit single-channel frames of one size")
    pts = np.asarray(prevPts, np.float32)
    shape = pts.shape                      # cv2 accepts (n, 2) and (n, 1, 2) and answers in kind
    pts = pts.reshape(-1, 2)
    init = None
    if flags & OPTFLOW_USE_INITIAL_FLOW:
        if nextPts is None or np.asarray(nextPts).size != pts.size:
            raise ValueError("OPTFLOW_USE_INITIAL_FLOW needs nextPts of prevPts' size")
        init = np.asarray(nextPts, np.float32).reshape(-1, 2)
    c = ctx or frontend.default_context(prev.shape[0], prev.shape[1])
    c.upload(0, prev)
    c.upload(1, nxt)
    out, st, er = c.klt_track(0, 1, pts, win=winSize, max_level=maxLevel, crit_type=criteria[0], max_count=criteria[1],
                              epsilon=criteria[2], flags=flags, min_eig=minEigThreshold, init_pts=init)
    return out.reshape(shape), st.reshape(-1, 1), er.reshape(-1, 1)
