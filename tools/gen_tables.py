#!/usr/bin/env python3
"""Dev-time table extraction (runs only in the build container, where /root/reference exists).

Extracts the *numeric data* (not code) that both the product's host-side table builder and the
oracle need:

  * cubature rules + degree->nIP map, as stored by the reference in
    src/element/Cubature.cpp:66-2730 (Witherden-Vincent rules and Gauss-Legendre points);
  * 1-D Gauss-Lobatto node lists of src/element/ReferenceElement.cpp:636-878.

It interprets the literal assignments of those two functions statement by statement (no C++
compiler needed) so that the resulting doubles are bit-identical to what the reference would hold
in memory, then cross-checks them against ressources/CubatureRules/expanded/*.txt.

Outputs
  oracle/tables/tables.json                      (oracle; hex-float strings => exact)
  hyperfox_b200/csrc/host/hfx_tables_data.inc    (product; C++ initialisers, %a hex floats)
"""
import json, math, os, re, sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _resize(lst, n, fill):
    if len(lst) > n:
        del lst[n:]
    while len(lst) < n:
        lst.append(fill())


def extract_cubature():
    src = open(os.path.join(REF, "src/element/Cubature.cpp")).read()
    body = src[src.index("void Cubature::initializeDatabase()"):src.index("};//initializeDatabase")]
    lines = body.split("\n")
    env = dict(tempDim=0, tempNumIPs=0, tempCoords=[], tempWeights=[], sqrt=math.sqrt, ceil=math.ceil)
    nip_map = {}     # (dim, degree, geom) -> nIP
    rules = {}       # (dim, nIP, geom) -> (coords, weights)
    # dimension 0 / order 0 / 1-D nIP map are computed, not tabulated (Cubature.cpp:80-112)
    vol_simplex = [0.0, 2.0, 2.0, 4.0 / 3.0]
    for d in range(1, 4):
        nip_map[(d, 0, "orthotope")] = 1
        rules[(d, 1, "orthotope")] = ([[0.0] * d], [float(2 ** d)])
        nip_map[(d, 0, "simplex")] = 1
        rules[(d, 1, "simplex")] = ([[-0.5] * d], [vol_simplex[d]])
    for j in range(1, 21):
        for g in ("orthotope", "simplex"):
            nip_map[(1, j, g)] = int(math.ceil((j + 1.0) / 2.0))
    i = 0
    started = False
    pvec = None
    while i < len(lines):
        ln = lines[i].strip()
        i += 1
        if ln.startswith("//dimension 1"):
            started = True
        if not started or not ln or ln.startswith("//"):
            continue
        if ln.startswith("for("):
            # two loop shapes only: resize loop (1 body line) and the pvec unpack loop
            if "tempCoords[i]" in lines[i] or "(tempCoords[i])" in lines[i]:
                for c in env["tempCoords"]:
                    _resize(c, env["tempDim"], float)
                i += 2
                continue
            if "nIPMap" in lines[i]:
                i += 3
                continue
            # unpack loop
            assert "for(int i = 0; i < tempDim" in lines[i], lines[i]
            idx = 0
            for j in range(env["tempNumIPs"]):
                for k in range(env["tempDim"]):
                    env["tempCoords"][j][k] = pvec[idx]; idx += 1
                env["tempWeights"][j] = pvec[idx]; idx += 1
            assert idx == len(pvec), (idx, len(pvec))
            while "}" not in lines[i] or lines[i].strip() != "}":
                i += 1
            # skip the closing of inner+outer loops: consume until 'delete tempPvec'
            while "delete tempPvec" not in lines[i]:
                i += 1
            i += 1
            continue
        if ln.startswith("tempPvec = new"):
            txt = ln[ln.index("("):] if "{" in ln else ""
            while "}" not in txt:
                txt += " " + lines[i]
                i += 1
            txt = txt[txt.index("{") + 1:txt.index("}")]
            pvec = [float(t) for t in txt.replace("\n", " ").split(",") if t.strip()]
            continue
        if ln.startswith("index = 0"):
            continue
        for st in [s.strip() for s in ln.split(";") if s.strip()]:
            if st in (")", "}", "})"):
                continue
            m = re.match(r"nIPMap\[cubDataKey\((\w+), (\w+), (\w+)\)\] = (\w+)$", st)
            if m:
                d = env["tempDim"] if m.group(1) == "tempDim" else int(m.group(1))
                v = env["tempNumIPs"] if m.group(4) == "tempNumIPs" else int(m.group(4))
                nip_map[(d, int(m.group(2)), m.group(3))] = v
                continue
            m = re.match(r"rulesDatabase\[cubDataKey\(tempDim, tempNumIPs, (\w+)\)\] = cubDataVal\(tempCoords, tempWeights\)$", st)
            if m:
                n = env["tempNumIPs"]
                assert len(env["tempCoords"]) == n and len(env["tempWeights"]) == n
                rules[(env["tempDim"], n, m.group(1))] = ([list(c) for c in env["tempCoords"]], list(env["tempWeights"]))
                continue
            if st == "tempCoords.resize(tempNumIPs)":
                _resize(env["tempCoords"], env["tempNumIPs"], list); continue
            if st == "tempWeights.resize(tempNumIPs)":
                _resize(env["tempWeights"], env["tempNumIPs"], float); continue
            if re.match(r"(tempDim|tempNumIPs) = \d+$", st) or re.match(r"tempCoords\[\d+\](\[\d+\])? = ", st) \
                    or re.match(r"tempWeights\[\d+\] = ", st):
                py = st.replace("std::sqrt", "sqrt")
                py = re.sub(r"std::vector<double>\(1, (.*)\)$", r"[float(\1)]", py)
                exec(py, env)
                continue
            raise RuntimeError("unhandled statement: " + st)
    return nip_map, rules


def extract_lobatto():
    src = open(os.path.join(REF, "src/element/ReferenceElement.cpp")).read()
    a = src.index("//orthotope and simplex dimension 1")
    b = src.index("std::vector<double> lobattoPts;")
    sec = src[a:b]
    out = {}
    for m in re.finditer(r"tempOrder = (\d+);.*?new std::vector<double>\(\s*\{(.*?)\}", sec, re.S):
        out[int(m.group(1))] = [float(t) for t in m.group(2).replace("\n", " ").split(",")]
    return out


def check_against_txt(rules):
    base = os.path.join(REF, "ressources/CubatureRules/expanded")
    names = {("simplex", 2): "tri", ("simplex", 3): "tet", ("orthotope", 2): "quad", ("orthotope", 3): "hex"}
    nchk = 0
    worst = 0.0
    for (d, n, g), (c, w) in rules.items():
        if (g, d) not in names:
            continue
        dn = os.path.join(base, names[(g, d)])
        fn = [f for f in os.listdir(dn) if f.endswith("-%d.txt" % n)]
        if not fn:
            continue
        for f in fn:
            rows = [[float(t) for t in l.split()] for l in open(os.path.join(dn, f)) if l.strip()]
            if len(rows) != n:
                continue
            err = max(abs(r[k] - (c[j] + [w[j]])[k]) for j, r in enumerate(rows) for k in range(d + 1))
            worst = max(worst, err)
            nchk += 1
    return nchk, worst


def main():
    nip_map, rules = extract_cubature()
    lob = extract_lobatto()
    nchk, worst = check_against_txt(rules)
    print("cubature rules: %d, nIP map entries: %d, txt cross-checks: %d (max |diff| %.3g), lobatto orders: %s"
          % (len(rules), len(nip_map), nchk, worst, sorted(lob)))
    # sanity: weights sum to the reference-domain volume
    for (d, n, g), (c, w) in rules.items():
        vol = {"simplex": [0, 2.0, 2.0, 4.0 / 3.0], "orthotope": [0, 2.0, 4.0, 8.0]}[g][d]
        assert abs(sum(w) - vol) < 1e-13, ((d, n, g), sum(w))
    js = {
        "_comment": "generated by tools/gen_tables.py from the reference's literal tables; hex floats are exact",
        "nip_map": {"%d,%d,%s" % k: v for k, v in sorted(nip_map.items())},
        "rules": {"%d,%d,%s" % k: {"coords": [[x.hex() for x in p] for p in c], "weights": [x.hex() for x in w]}
                  for k, (c, w) in sorted(rules.items())},
        "lobatto": {str(k): [x.hex() for x in v] for k, v in sorted(lob.items())},
    }
    os.makedirs(os.path.join(ROOT, "oracle/tables"), exist_ok=True)
    json.dump(js, open(os.path.join(ROOT, "oracle/tables/tables.json"), "w"), indent=0)
    # C++ initialisers for the product
    geo = {"simplex": 0, "orthotope": 1}
    o = ["// GENERATED by tools/gen_tables.py -- numeric tables only (cubature rules, 1-D Lobatto nodes).",
         "// Source of the numbers: reference src/element/Cubature.cpp:66-2730, ReferenceElement.cpp:636-878.",
         "struct HfxNipEntry { int dim, degree, geom, nip; };",
         "static const HfxNipEntry kHfxNipMap[] = {"]
    for (d, deg, g), v in sorted(nip_map.items()):
        o.append("  {%d, %d, %d, %d}," % (d, deg, geo[g], v))
    o.append("};")
    o.append("struct HfxRuleEntry { int dim, nip, geom, offset; };")
    flat = []
    ent = []
    for (d, n, g), (c, w) in sorted(rules.items()):
        ent.append("  {%d, %d, %d, %d}," % (d, n, geo[g], len(flat)))
        for j in range(n):
            flat.extend(c[j]); flat.append(w[j])
    o.append("static const HfxRuleEntry kHfxRules[] = {")
    o.extend(ent)
    o.append("};")
    o.append("// per rule: nip rows of (dim coords, weight)")
    o.append("static const double kHfxRuleData[] = {")
    for k in range(0, len(flat), 4):
        o.append("  " + ", ".join(x.hex() for x in flat[k:k + 4]) + ",")
    o.append("};")
    o.append("struct HfxLobattoEntry { int order, offset; };")
    lflat = []
    o.append("static const HfxLobattoEntry kHfxLobatto[] = {")
    for k, v in sorted(lob.items()):
        o.append("  {%d, %d}," % (k, len(lflat)))
        lflat.extend(v)
    o.append("};")
    o.append("static const double kHfxLobattoData[] = {")
    for k in range(0, len(lflat), 4):
        o.append("  " + ", ".join(x.hex() for x in lflat[k:k + 4]) + ",")
    o.append("};")
    os.makedirs(os.path.join(ROOT, "hyperfox_b200/csrc/host"), exist_ok=True)
    open(os.path.join(ROOT, "hyperfox_b200/csrc/host/hfx_tables_data.inc"), "w").write("\n".join(o) + "\n")


if __name__ == "__main__":
    main()
