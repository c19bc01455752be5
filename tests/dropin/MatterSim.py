"""Stand-in for the compiled ``MatterSim`` module — TEST INFRASTRUCTURE ONLY (tests/test_dropin_gpu.py).

The reference's simulator is C++/OpenGL and cannot be built here.  The agents only use its
navigation-graph side with rendering disabled (env.py:241-247), so this module restates exactly that from the
published behaviour of the simulator: the discretised camera (12 headings x 3 elevations, view index = 12 * elevation
level + heading step, src/lib/MatterSim.cpp:339-367), ``newEpisode`` / ``makeAction`` (:379-435, :470-508) and the
navigable-location rule (:276-311: unobstructed + included neighbours inside the horizontal field of view, sorted by
angular distance from the view centre, the current location first), over the reference's own
``connectivity/<scan>_connectivity.json`` files read relative to the working directory like utils.py:37 does.
"""
import json
import math


class _Loc(object):
    __slots__ = ("viewpointId", "ix", "point", "rel_heading", "rel_elevation", "rel_distance")

    def __init__(self, viewpointId, ix, point, rel_heading=0.0, rel_elevation=0.0, rel_distance=0.0):
        self.viewpointId, self.ix, self.point = viewpointId, ix, point
        self.rel_heading, self.rel_elevation, self.rel_distance = rel_heading, rel_elevation, rel_distance


class _State(object):
    def __init__(self):
        self.scanId = ""
        self.step = 0
        self.rgb = None
        self.location = None
        self.heading = 0.0
        self.elevation = 0.0
        self.viewIndex = 0
        self.navigableLocations = []


_GRAPHS = {}


def _load(scan):
    if scan not in _GRAPHS:
        with open("connectivity/%s_connectivity.json" % scan) as f:
            data = json.load(f)
        locs = []
        for item in data:
            p = item["pose"]
            locs.append((item["image_id"], bool(item["included"]), (p[3], p[7], p[11]), [bool(u) for u in item["unobstructed"]]))
        _GRAPHS[scan] = locs
    return _GRAPHS[scan]


class Simulator(object):
    HEADINGS = 12
    ELEV_INC = math.pi / 6.0

    def __init__(self):
        self.state = _State()
        self.width, self.height, self.vfov = 320, 240, 0.8
        self.discretize = False
        self.initialized = False

    # configuration calls of env.py:241-247
    def setRenderingEnabled(self, flag): pass
    def setDiscretizedViewingAngles(self, flag): self.discretize = bool(flag)
    def setCameraResolution(self, w, h): self.width, self.height = int(w), int(h)
    def setCameraVFOV(self, vfov): self.vfov = float(vfov)
    def setNavGraphPath(self, path): pass
    def setDatasetPath(self, path): pass
    def init(self): self.initialized = True
    def close(self): pass

    def _set_heading_elevation(self, heading, elevation):
        st = self.state
        heading = math.fmod(heading, 2.0 * math.pi)
        while heading < 0.0:
            heading += 2.0 * math.pi
        st.heading = heading
        if self.discretize:
            inc = 2.0 * math.pi / self.HEADINGS
            step = int(math.floor(heading / inc + 0.5))          # lround of a non-negative value
            if step == self.HEADINGS:
                step = 0
            st.heading = step * inc
            if elevation < -self.ELEV_INC / 2.0:
                st.elevation, st.viewIndex = -self.ELEV_INC, step
            elif elevation > self.ELEV_INC / 2.0:
                st.elevation, st.viewIndex = self.ELEV_INC, step + 2 * self.HEADINGS
            else:
                st.elevation, st.viewIndex = 0.0, step + self.HEADINGS
        else:
            st.elevation = max(min(elevation, math.pi / 2 - 0.01), -math.pi / 2 + 0.01)

    def _populate(self):
        st = self.state
        locs = _load(st.scanId)
        idx = st.location.ix
        cur = locs[idx][2]
        adj = math.pi / 2.0 - st.heading
        hx, hy = math.cos(adj), math.sin(adj)
        cos_half_hfov = math.cos(self.vfov * self.width / self.height / 2.0)
        out = []
        for i, (vid, included, pos, _) in enumerate(locs):
            if i == idx or not (locs[idx][3][i] and included):
                continue
            dx, dy, dz = pos[0] - cur[0], pos[1] - cur[1], pos[2] - cur[2]
            dist = math.sqrt(dx * dx + dy * dy + dz * dz)
            planar = math.sqrt(dx * dx + dy * dy)
            rel_el = math.atan2(dz, planar) - st.elevation
            if planar == 0.0:
                continue
            cos_angle = (dx * hx + dy * hy) / planar
            if cos_angle >= cos_half_hfov:
                rel_h = math.atan2(dx * hy - dy * hx, dx * hx + dy * hy)
                out.append(_Loc(vid, i, pos, rel_h, rel_el, dist))
        out.sort(key=lambda l: math.sqrt(l.rel_heading ** 2 + l.rel_elevation ** 2))
        st.navigableLocations = [st.location] + out

    def newEpisode(self, scanId, viewpointId, heading=0.0, elevation=0.0):
        st = self.state
        st.step = 0
        self._set_heading_elevation(heading, elevation)
        st.scanId = scanId
        locs = _load(scanId)
        ix = -1
        for i, (vid, included, pos, _) in enumerate(locs):
            if vid == viewpointId:
                if not included:
                    raise ValueError("MatterSim: ViewpointId: %s, is excluded from the connectivity graph." % viewpointId)
                ix = i
                break
        if ix < 0:
            raise ValueError("MatterSim: Could not find viewpointId: %s, is viewpoint id valid?" % viewpointId)
        st.location = _Loc(locs[ix][0], ix, locs[ix][2])
        self._populate()

    def getState(self):
        return self.state

    def makeAction(self, index, heading, elevation):
        st = self.state
        if index < 0 or index >= len(st.navigableLocations):
            raise IndexError("MatterSim: Invalid action index: %d" % index)
        nxt = st.navigableLocations[index]
        st.location = _Loc(nxt.viewpointId, nxt.ix, nxt.point)
        st.step += 1
        if self.discretize:
            inc = 2.0 * math.pi / self.HEADINGS
            heading = inc if heading > 0 else (-inc if heading < 0 else 0.0)
            elevation = self.ELEV_INC if elevation > 0 else (-self.ELEV_INC if elevation < 0 else 0.0)
        self._set_heading_elevation(st.heading + heading, st.elevation + elevation)
        self._populate()
