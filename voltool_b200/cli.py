"""Tcl-free command line over the C ABI: ``python -m voltool_b200 <command> [options]``.

Command and option names follow the reference's dispatch table (TclVoltool.C:3650-3703) and usage texts
(e.g. TclVoltool.C:1079-1084, 2650-2661; TclMDFF.C:168-184).  Where the reference takes a VMD molecule
(``-mol/-vol``) or an atom selection, this front end takes files instead: maps are OpenDX (``-i/-i1/-i2``,
``-o``), atoms are a text table ``x y z radius [mass]`` per line (``-atoms``).  Tcl lists such as
``-amt {1 2 3 4 5 6}`` may be written quoted (``-amt "1 2 3 4 5 6"``) or as separate words.
Everything runs on the GPU library; there is no CPU fallback."""
import sys

import numpy as np

FLT_MAX = 3.4028234663852886e38

USAGE = """usage: python -m voltool_b200 <command> [options]
one-map commands (-i <map.dx> [-o <out.dx>]):
  com | info [origin|xsize|ysize|zsize|minmax] | hist [-nbins n] | sigma | downsample | supersample
  smooth -sigma x | smult -amt x | sadd -amt x | range -minmax {min max} | clamp [-min a] [-max b]
  binmask [-threshold x] | pot [-threshold x] | trim -amt {x1 x2 y1 y2 z1 z2}
  crop -amt {minx miny minz maxx maxy maxz} | moveto -pos {x y z} | move -mat {16 numbers, row major}
two-map commands (-i1 <a.dx> -i2 <b.dx>):
  add | diff | mult | avg  [-union] [-nointerp] [-o <out.dx>]
  correlate [-thresholddensity x]
atom commands (-atoms <table: x y z radius [mass]> -res R):
  sim [-spacing g] [-o <out.dx>]
  cc  -i <target.dx> [-spacing g] [-thresholddensity x] [-savesynthmap <out.dx>] [-savediffmap <out.dx>]
      [-savespatialccmap <out.dx>]
  fit -i <target.dx> [-o <fitted atoms table>]
"""

# how many values each option takes (0 = flag); list options accept one quoted word or that many words
NARGS = {"-i": 1, "-o": 1, "-i1": 1, "-i2": 1, "-atoms": 1, "-res": 1, "-spacing": 1, "-sigma": 1, "-amt": -1,
         "-minmax": 2, "-min": 1, "-max": 1, "-threshold": 1, "-nbins": 1, "-pos": 3, "-mat": 16,
         "-thresholddensity": 1, "-savesynthmap": 1, "-savediffmap": 1, "-savespatialccmap": 1, "-union": 0, "-nointerp": 0}
AMT_LEN = {"trim": 6, "crop": 6, "smult": 1, "sadd": 1}


class UsageError(Exception):
    pass


def parse(cmd, argv):
    opts, pos = {}, []
    i = 0
    while i < len(argv):
        a = argv[i]
        if a not in NARGS:
            if a.startswith("-") and not _is_number(a):
                raise UsageError(f"unknown option {a}")
            pos.append(a)
            i += 1
            continue
        n = NARGS[a]
        if n == -1:
            n = AMT_LEN.get(cmd, 1)
        if n == 0:
            opts[a] = True
            i += 1
            continue
        if i + 1 >= len(argv):
            raise UsageError(f"{a} needs a value")
        first = argv[i + 1].strip("{}").split()
        if n == 1 or len(first) == n:
            vals, i = (first if n > 1 else [argv[i + 1]]), i + 2
        else:
            vals, i = [v.strip("{}") for v in argv[i + 1:i + 1 + n]], i + 1 + n
            vals = [v for v in vals if v]
        if len(vals) != n:
            raise UsageError(f"{a} needs {n} value(s)")
        opts[a] = vals if n > 1 else vals[0]
    return opts, pos


def _is_number(s):
    try:
        float(s)
        return True
    except ValueError:
        return False


def _floats(v):
    return [float(x) for x in (v if isinstance(v, list) else [v])]


def read_atoms(path):
    t = np.loadtxt(path, dtype=np.float64, ndmin=2)
    if t.shape[1] < 4:
        raise UsageError(f"{path}: need at least 4 columns (x y z radius [mass])")
    pos = np.ascontiguousarray(t[:, :3], np.float32)
    radii = np.ascontiguousarray(t[:, 3], np.float32)
    mass = np.ascontiguousarray(t[:, 4] if t.shape[1] > 4 else np.ones(len(t)), np.float32)
    return pos, radii, mass


def write_atoms(path, pos, radii, mass):
    np.savetxt(path, np.column_stack([pos.reshape(-1, 3), radii, mass]), fmt="%.9g")


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help", "help"):
        print(USAGE)
        return 0 if argv else 1
    cmd, rest = argv[0].lower(), argv[1:]
    try:
        opts, pos = parse(cmd, rest)
        return run(cmd, opts, pos)
    except UsageError as e:
        print(f"voltool {cmd}: {e}\n\n{USAGE}", file=sys.stderr)
        return 1


def run(cmd, o, pos):
    import voltool_b200 as vt
    vt.init()

    def need(k):
        if k not in o:
            raise UsageError(f"{k} is required")
        return o[k]

    def load(k="-i"):
        return vt.VolMap.read_file(need(k))      # .mrc / .map / .ccp4 or OpenDX

    def save(m, k="-o"):
        if k in o:
            m.write_file(o[k])

    one_map = {"smooth", "smult", "sadd", "range", "clamp", "binmask", "pot", "sigma", "downsample", "supersample",
               "trim", "crop", "moveto", "move"}
    if cmd in one_map:
        m = load()
        if cmd == "smooth":
            m.gaussian_blur(float(need("-sigma")))                       # TclVoltool.C:1643
        elif cmd == "smult":
            m.scale_by(_floats(need("-amt"))[0])                         # :1535
        elif cmd == "sadd":
            m.scalar_add(_floats(need("-amt"))[0])                       # :1751
        elif cmd == "range":
            lo, hi = _floats(need("-minmax"))
            m.rescale_voxel_value_range(lo, hi)                          # :1873
        elif cmd == "clamp":
            mn, mx = m.datarange()                                       # defaults: the existing range (usage :1323-1324)
            m.clamp(float(o.get("-min", mn)), float(o.get("-max", mx)))  # :1427
        elif cmd == "binmask":
            m.binmask(float(o.get("-threshold", 0.0)))                   # :2635, default 0 (:2552)
        elif cmd == "pot":
            m.mdff_potential(float(o.get("-threshold", 0.0)))
        elif cmd == "sigma":
            m.sigma_scale()                                              # :2167
        elif cmd == "downsample":
            m.downsample()                                               # :1971
        elif cmd == "supersample":
            m.supersample()                                              # :2069
        elif cmd == "trim":
            t = [int(v) for v in _floats(need("-amt"))]
            m.pad(*[-v for v in t])                                      # trim = pad(-t...), :1185
        elif cmd == "crop":
            m.crop(*_floats(need("-amt")))                               # :1306
        elif cmd == "moveto":
            m.vol_moveto(m.vol_com(), _floats(need("-pos")))             # Voltool.C:379-391
        elif cmd == "move":
            rm = np.array(_floats(need("-mat")), np.float32).reshape(4, 4)
            m.vol_move(np.ascontiguousarray(rm.T).ravel())               # Matrix4 is column major
        save(m)
        return 0
    if cmd == "com":
        print("%g %g %g" % tuple(load().vol_com()))
        return 0
    if cmd == "hist":
        bins, mid = load().histogram(int(o.get("-nbins", 10)))           # default 10 bins (:2304)
        print(" ".join(str(int(b)) for b in bins))
        print(" ".join("%g" % v for v in mid))
        return 0
    if cmd == "info":
        m = load()
        g = m.grid
        what = pos[0] if pos else None
        if what in (None, "origin"):
            print("origin %g %g %g" % tuple(g.origin))
        if what in (None, "xsize"):
            print("xsize %d" % g.xsize)
        if what in (None, "ysize"):
            print("ysize %d" % g.ysize)
        if what in (None, "zsize"):
            print("zsize %d" % g.zsize)
        if what in (None, "minmax"):
            print("minmax %g %g" % m.datarange())
        return 0
    if cmd in ("add", "diff", "mult", "avg"):
        a, b = load("-i1"), load("-i2")
        op = {"add": vt.add, "diff": vt.subtract, "mult": vt.multiply, "avg": vt.average}[cmd]
        save(op(a, b, interp="-nointerp" not in o, use_union="-union" in o))   # defaults :2674-2675
        return 0
    if cmd == "correlate":
        cc = vt.cc_threaded(load("-i1"), load("-i2"), float(o.get("-thresholddensity", -FLT_MAX)))   # :3376, :3505
        print("%.9g" % cc)
        return 0
    if cmd in ("sim", "cc", "fit"):
        apos, radii, mass = read_atoms(need("-atoms"))
        res = float(need("-res"))
        gs = float(o.get("-spacing", 0.0))
        if cmd == "sim":
            radscale, gspacing, quality = vt.sim_policy(res, gs)         # TclMDFF.C:123-135
            save(vt.sim(apos, radii, quality, radscale, gspacing))
            return 0
        target = load()
        if cmd == "cc":
            thr = float(o.get("-thresholddensity", -FLT_MAX))            # TclMDFF.C:216
            if "-savediffmap" in o or "-savespatialccmap" in o:           # -calcdiffmap / -calcspatialccmap, :180-184
                cc, synth, diff, spat = vt.sim_cc_maps(apos, radii, res, target, thr, gs)
                for key, m in (("-savesynthmap", synth), ("-savediffmap", diff), ("-savespatialccmap", spat)):
                    if key in o:
                        m.write_file(o[key])
                print("%.9g" % cc)
                return 0
            want = "-savesynthmap" in o
            r = vt.sim_cc(apos, radii, res, target, thr, gs, want_synth=want)
            if want:
                cc, synth = r
                synth.write_file(o["-savesynthmap"])
            else:
                cc = r
            print("%.9g" % cc)
            return 0
        newpos, result = vt.fit(apos, radii, mass, target, res)          # TclVoltool.C:784-916
        if "-o" in o:
            write_atoms(o["-o"], newpos, radii, mass)
        print("Best CC:%f" % result.best_cc)                             # :926
        return 0
    raise UsageError("unknown command")


if __name__ == "__main__":
    sys.exit(main())
