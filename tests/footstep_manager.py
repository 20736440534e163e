"""Test fixture: footstep queue -> reference ZMP and ZMP limits, restated from reference
tests/src/FootstepManager.h:19-254 (numpy, plain python)."""
import bisect

import numpy as np

LEFT, RIGHT = 0, 1
EPS_T = 1e-6


def opposite(foot):
    return RIGHT if foot == LEFT else LEFT


class Footstep:
    def __init__(self, foot, pos, transit_start_time, transit_duration, swing_duration):  # :66-77
        self.foot, self.pos = foot, np.asarray(pos, dtype=np.float64)
        self.transit_start_time = transit_start_time
        self.swing_start_time = transit_start_time + 0.5 * transit_duration
        self.swing_end_time = transit_start_time + 0.5 * transit_duration + swing_duration
        self.transit_end_time = transit_start_time + transit_duration + swing_duration


def mid_pos(stance):  # :89-104
    if len(stance) == 1:
        return next(iter(stance.values())).copy()
    return 0.5 * (stance[LEFT] + stance[RIGHT])


def support_region(stance):  # :107-129
    if len(stance) == 1:
        p = next(iter(stance.values()))
        return p.copy(), p.copy()
    return np.minimum(stance[LEFT], stance[RIGHT]), np.maximum(stance[LEFT], stance[RIGHT])


class FootstepManager:
    def __init__(self, initial_footstance=None):
        self.footstance = initial_footstance or {LEFT: np.array([0.0, 0.1]), RIGHT: np.array([0.0, -0.1])}
        self.footstep_list = []
        self.horizon_duration = 10.0
        self.foot_size = np.array([0.1, 0.05])
        self.ref_zmp_list, self.ref_footstance_list = {}, {}
        self.prev_footstep = None

    def append_footstep(self, fs):  # :217-226
        if self.footstep_list and fs.transit_start_time < self.footstep_list[-1].transit_end_time:
            raise RuntimeError("transit_start_time of specified footstep must be after transit_end_time of last footstep")
        self.footstep_list.append(fs)

    def update(self, t):  # :147-212
        fl = self.footstep_list
        if fl and fl[0].swing_end_time <= t:
            self.footstance[fl[0].foot] = fl[0].pos
        while fl and fl[0].transit_end_time < t:
            self.prev_footstep = fl.pop(0)
        zl, sl = {}, {}

        def emplace(d, k, v):  # std::map::emplace keeps the first value of a key
            d.setdefault(k, v)

        if not fl:
            emplace(zl, t, mid_pos(self.footstance))
            emplace(zl, t + self.horizon_duration, mid_pos(self.footstance))
            emplace(sl, t, dict(self.footstance))
            emplace(sl, t + self.horizon_duration, dict(self.footstance))
        else:
            if t < fl[0].transit_start_time:
                emplace(zl, t, mid_pos(self.footstance))
                emplace(sl, t, dict(self.footstance))
            tmp = dict(self.footstance)
            for fs in fl:
                if fs.transit_start_time > t + self.horizon_duration:
                    break
                emplace(zl, fs.transit_start_time, mid_pos(tmp))
                emplace(sl, fs.transit_start_time, dict(tmp))
                tmp.pop(fs.foot, None)
                emplace(zl, fs.swing_start_time, tmp[opposite(fs.foot)].copy())
                emplace(sl, fs.swing_start_time, dict(tmp))
                tmp[fs.foot] = fs.pos
                emplace(zl, fs.swing_end_time, tmp[opposite(fs.foot)].copy())
                emplace(sl, fs.swing_end_time, dict(tmp))
                emplace(zl, fs.transit_end_time, mid_pos(tmp))
            if max(zl) < t + self.horizon_duration:
                emplace(zl, t + self.horizon_duration, mid_pos(tmp))
                emplace(sl, t + self.horizon_duration, dict(tmp))
        self.ref_zmp_list, self.ref_footstance_list = zl, sl
        self._zk, self._sk = sorted(zl), sorted(sl)

    def ref_zmp(self, t):  # :228-237
        t += EPS_T
        i = bisect.bisect_right(self._zk, t)
        t0, t1 = self._zk[i - 1], self._zk[i]
        ratio = (t - t0) / (t1 - t0)
        return (1 - ratio) * self.ref_zmp_list[t0] + ratio * self.ref_zmp_list[t1]

    def zmp_limits(self, t):  # :242-254
        t += EPS_T
        i = bisect.bisect_right(self._sk, t)
        lo, hi = support_region(self.ref_footstance_list[self._sk[i - 1]])
        return lo - 0.5 * self.foot_size, hi + 0.5 * self.foot_size

    def make_linear_mpc_zmp_ref_data(self, t):  # :356-365 (the epsilon is added twice, as in the reference)
        return self.zmp_limits(t + EPS_T)

    def make_ismpc_ref_data(self, t):  # :370-380
        t += EPS_T
        return self.ref_zmp(t), self.zmp_limits(t)


def make_dcm_tracking_ref_data(fm, current_time, horizon_duration=5.0):
    """FootstepManager::makeDcmTrackingRefData (:260-274; the reference declares horizon_duration as int) ->
    (current_zmp[2], [(time, zmp[2]), ...])."""
    t = current_time + EPS_T
    i = bisect.bisect_right(fm._zk, t)
    current_zmp = fm.ref_zmp_list[fm._zk[i - 1]]
    knots = []
    while i < len(fm._zk) and fm._zk[i] < t + int(horizon_duration):
        knots.append((fm._zk[i], fm.ref_zmp_list[fm._zk[i]]))
        i += 1
    return current_zmp, knots


def make_foot_guided_control_ref_data(fm, current_time):
    """FootstepManager::makeFootGuidedControlRefData (:279-351) -> dict(transit_start_zmp, transit_end_zmp,
    transit_start_time, transit_duration)."""
    t = current_time + EPS_T
    constant_zmp_duration, concat_thre, horizon_margin = 1.0, 0.1, 1e-3
    fl = fm.footstep_list
    if not fl:
        mid = mid_pos(fm.footstance)
        return dict(transit_start_zmp=mid, transit_end_zmp=mid.copy(), transit_start_time=t + constant_zmp_duration, transit_duration=0.0)
    fs = fl[0]
    if t < fs.swing_start_time:
        rd = dict(transit_start_zmp=mid_pos(fm.footstance), transit_end_zmp=fm.footstance[opposite(fs.foot)].copy(),
                  transit_start_time=fs.transit_start_time, transit_duration=fs.swing_start_time - fs.transit_start_time)
        prev = fm.prev_footstep
        if prev is not None and prev.foot == opposite(fs.foot) and fs.transit_start_time - prev.transit_end_time < concat_thre:
            rd["transit_start_zmp"] = fm.footstance[opposite(prev.foot)].copy()
            rd["transit_start_time"] = prev.swing_end_time
            rd["transit_duration"] = fs.swing_start_time - prev.swing_end_time
    else:
        tmp = dict(fm.footstance)
        tmp[fs.foot] = fs.pos
        rd = dict(transit_start_zmp=fm.footstance[opposite(fs.foot)].copy(), transit_end_zmp=mid_pos(tmp),
                  transit_start_time=fs.swing_end_time, transit_duration=fs.transit_end_time - fs.swing_end_time)
        if len(fl) >= 2:
            nxt = fl[1]
            if nxt.foot == opposite(fs.foot) and nxt.transit_start_time - fs.transit_end_time < concat_thre:
                rd["transit_end_zmp"] = fs.pos.copy()
                rd["transit_duration"] = nxt.swing_start_time - fs.swing_end_time
    if rd["transit_start_time"] + rd["transit_duration"] < t + horizon_margin:
        rd["transit_duration"] += horizon_margin
    return rd


def walking_plan(step_length=0.2, step_width=0.2, transit_duration=0.2, swing_duration=0.8):
    """The six-step plan of reference tests/src/TestLinearMpcZmp.cpp:30-41 (defaults reproduce it)."""
    m = FootstepManager({LEFT: np.array([0.0, 0.5 * step_width]), RIGHT: np.array([0.0, -0.5 * step_width])})
    L, w = step_length, 0.5 * step_width
    for foot, x, t0 in [(LEFT, L, 2.0), (RIGHT, 2 * L, 3.0), (LEFT, 3 * L, 4.0), (RIGHT, 4 * L, 5.0), (LEFT, 3 * L, 6.0),
                        (RIGHT, 3 * L, 7.0)]:
        m.append_footstep(Footstep(foot, (x, w if foot == LEFT else -w), t0, transit_duration, swing_duration))
    return m


def make_step_mpc_ref_data(fm, current_time):
    """FootstepManager::makeStepMpcRefData (:385-448) -> [(is_single_support, zmp[2], end_time), ...]."""
    t = current_time - EPS_T
    constant_zmp_duration, horizon_duration = 0.2, 3.0
    fl = fm.footstep_list
    if not fl:
        return [(False, mid_pos(fm.footstance), t + constant_zmp_duration)]
    elements = []
    tmp = {k: v.copy() for k, v in fm.footstance.items()}
    if t < fl[0].swing_start_time:
        elements.append((False, mid_pos(tmp), fl[0].swing_start_time))
    for i, fs in enumerate(fl):
        if i > 0 or t < fs.swing_end_time:
            elements.append((True, tmp[opposite(fs.foot)].copy(), fs.swing_end_time))
            tmp[fs.foot] = fs.pos.copy()
        elements.append((False, mid_pos(tmp), fs.transit_end_time if i == len(fl) - 1 else fl[i + 1].swing_start_time))
        if elements[-1][2] > t + horizon_duration:
            break
    return elements
