"""Synthetic hospital-schema tables (BASELINE.json configs[3]).

N Records over H hospitals (Zipf(1.0)), P places, C counties, S states, T hospital types,
M measures, 8 conditions; every observed cell independently corrupted with probability
`typo_rate` by exactly one uniformly chosen {insert, delete, substitute, transpose} with a–z
letters (mirrors `perform_typo`, reference add_typos.jl:9-32); no missing cells.  Option
lists are the unique dirty values per column, as `experiments/hospital/load_data.jl:18-19`
builds them.  Also returns the ground-truth latent tables, used as the initial trace.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from . import model as M
from ..lowering import VALUE_DTYPE, VAL_ABSENT, VAL_KEY, VAL_STR, FlatIR

LETTERS = "abcdefghijklmnopqrstuvwxyz"
DIGITS = "0123456789"


def _rand_words(rng, n: int, mean_len: int, lo: int, hi: int, alphabet: str = LETTERS + " ") -> List[str]:
    out, seen = [], set()
    while len(out) < n:
        L = int(np.clip(round(rng.normal(mean_len, max(1.0, mean_len / 4))), lo, hi))
        chars = rng.integers(0, len(alphabet), size=L)
        s = "".join(alphabet[c] for c in chars).strip()
        if len(s) < lo:
            s = s + "x" * (lo - len(s))
        if s not in seen:
            seen.add(s)
            out.append(s)
    return out


def _typo(rng, w: str) -> str:
    t = rng.integers(0, 4)
    L = len(w)
    letter = LETTERS[rng.integers(0, 26)]
    if t == 0:
        i = rng.integers(0, L + 1)
        return w[:i] + letter + w[i:]
    if t == 1 and L > 1:
        i = rng.integers(0, L)
        return w[:i] + w[i + 1:]
    if t == 2 and L > 1:
        i = rng.integers(0, L - 1)
        return w[:i] + w[i + 1] + w[i] + w[i + 2:]
    i = rng.integers(0, L)
    return w[:i] + letter + w[i + 1:]


def generate_hospital(n_rows: int, seed: int = 20260924, H: int = 4096, P: int = 2048, C: int = 512, S: int = 50,
                      T: int = 8, Mm: int = 64, n_cond: int = 8, typo_rate: float = 0.05):
    """Returns (dirty columns, truth) where truth holds the clean entities and per-row indices."""
    rng = np.random.default_rng(seed)
    H, P, C = min(H, max(4, n_rows)), min(P, max(4, n_rows)), min(C, max(4, n_rows))
    states = _rand_words(rng, S, 2, 2, 2, LETTERS)
    county_name = _rand_words(rng, C, 7, 3, 30)
    county_state = rng.integers(0, S, size=C)
    city = _rand_words(rng, P, 8, 3, 30)
    place_county = rng.integers(0, C, size=P)
    types = _rand_words(rng, T, 20, 10, 30)
    owners = _rand_words(rng, 10, 27, 5, 40)
    services = ["yes", "no"]
    provider = _rand_words(rng, H, 5, 5, 5, DIGITS)
    name = _rand_words(rng, H, 26, 3, 50)
    addr = _rand_words(rng, H, 20, 10, 30, LETTERS + DIGITS + " ")
    phone = _rand_words(rng, H, 10, 10, 10, DIGITS)
    zipc = _rand_words(rng, H, 5, 5, 5, DIGITS)
    h_owner = rng.integers(0, len(owners), size=H)
    h_service = rng.integers(0, 2, size=H)
    h_type = rng.integers(0, T, size=H)
    h_place = rng.integers(0, P, size=H)
    conds = _rand_words(rng, n_cond, 17, 5, 35)
    code = _rand_words(rng, Mm, 6, 4, 11, LETTERS + "-")
    mname = _rand_words(rng, Mm, 90, 46, 184)
    m_cond = rng.integers(0, n_cond, size=Mm)

    # drop entities nobody references (the reference garbage-collects zero-count rows)
    def compact(idx, *tables):
        used, inv = np.unique(idx, return_inverse=True)
        return inv, [([t[i] for i in used] if isinstance(t, list) else t[used]) for t in tables]
    h_place, (city, place_county) = compact(h_place, city, place_county)
    place_county, (county_name, county_state) = compact(place_county, county_name, county_state)
    h_type, (types,) = compact(h_type, types)
    m_cond, (conds,) = compact(m_cond, conds)
    P, C, T = len(city), len(county_name), len(types)

    w = 1.0 / np.arange(1, H + 1)
    row_h = rng.choice(H, size=n_rows, p=w / w.sum())
    row_m = rng.integers(0, Mm, size=n_rows)
    # every entity must be referenced at least once (no zero-count rows in the initial trace)
    if n_rows >= H:
        row_h[:H] = rng.permutation(H)
    if n_rows >= Mm:
        row_m[:Mm] = rng.permutation(Mm)

    def gather(values: List[str], idx: np.ndarray) -> List[str]:
        arr = np.array(values, dtype=object)
        return list(arr[idx])

    rp = h_place[row_h]
    rc = place_county[rp]
    clean_cols = {
        "ProviderNumber": gather(provider, row_h), "HospitalName": gather(name, row_h),
        "Address1": gather(addr, row_h), "City": gather(city, rp),
        "State": gather(states, county_state[rc]), "ZipCode": gather(zipc, row_h),
        "CountyName": gather(county_name, rc), "PhoneNumber": gather(phone, row_h),
        "HospitalType": gather(types, h_type[row_h]), "HospitalOwner": gather(owners, h_owner[row_h]),
        "EmergencyService": gather(services, h_service[row_h]), "Condition": gather(conds, m_cond[row_m]),
        "MeasureCode": gather(code, row_m), "MeasureName": gather(mname, row_m),
    }
    st = clean_cols["State"]; cd = clean_cols["MeasureCode"]
    clean_cols["Stateavg"] = [f"{a}_{b}" for a, b in zip(st, cd)]
    dirty: Dict[str, List[str]] = {}
    for col, vals in clean_cols.items():
        out = list(vals)
        hit = np.nonzero(rng.random(n_rows) < typo_rate)[0]
        # the first max(H, M) rows stay clean so that every entity's clean value is an option
        # of its column (ChooseProportionally.incorporate_choice! needs it; choose_proportionally.jl:57-60)
        hit = hit[hit >= max(H, Mm)] if n_rows > 2 * max(H, Mm) else hit[:0]
        for i in hit:
            out[i] = _typo(rng, out[i])
        dirty[col] = out
    # the clean value of every entity must itself be an option of the column
    # (ChooseProportionally.incorporate_choice! requires it; choose_proportionally.jl:57-60)
    truth = dict(states=states, county_name=county_name, county_state=county_state, city=city, place_county=place_county,
                 types=types, owners=owners, services=services, provider=provider, name=name, addr=addr, phone=phone,
                 zip=zipc, h_owner=h_owner, h_service=h_service, h_type=h_type, h_place=h_place, conds=conds, code=code,
                 mname=mname, m_cond=m_cond, row_h=row_h, row_m=row_m, clean=clean_cols)
    return dirty, truth


def truth_snapshot(model: M.PCleanModel, ir: FlatIR, truth: dict) -> dict:
    """Ground-truth latent tables + assignment in the form `load_trace_from_snapshot` takes
    (denormalised rows, keys 1..n per class)."""
    tables: Dict[str, np.ndarray] = {}

    def blank(cls: str, n: int) -> np.ndarray:
        cm = model.classes[cls]
        n_normal = sum(1 for x in cm.nodes if not isinstance(x, M.ExternalLikelihoodNode))
        a = np.zeros((n_normal, n), dtype=VALUE_DTYPE)
        a["tag"] = VAL_ABSENT
        return a

    def put_str(arr, cls, name, values: List[str]):
        v = model.classes[cls].names[name] - 1
        ids = np.array([ir.intern_string(s) for s in values], dtype=np.int32)
        arr[v]["tag"] = VAL_STR
        arr[v]["i"] = ids

    def put_fk(arr, cls, name, target: str, idx: np.ndarray):
        cm = model.classes[cls]
        v = cm.names[name]
        fk = cm.node(v)
        arr[v - 1]["tag"] = VAL_KEY
        arr[v - 1]["d"] = (idx + 1).astype(np.float64)
        tgt = tables[target]
        for tv, lv in fk.vmap.items():
            arr[lv - 1] = tgt[tv - 1][idx]

    t = truth
    C = len(t["county_name"]); P = len(t["city"]); H = len(t["name"]); Mm = len(t["code"])
    a = blank("County", C)
    put_str(a, "County", "state", [t["states"][i] for i in t["county_state"]])
    put_str(a, "County", "county", t["county_name"])
    tables["County"] = a
    a = blank("Place", P)
    put_fk(a, "Place", "county", "County", t["place_county"])
    put_str(a, "Place", "city", t["city"])
    tables["Place"] = a
    a = blank("Condition", len(t["conds"]))
    put_str(a, "Condition", "desc", t["conds"])
    tables["Condition"] = a
    a = blank("Measure", Mm)
    put_str(a, "Measure", "code", t["code"])
    put_str(a, "Measure", "name", t["mname"])
    put_fk(a, "Measure", "condition", "Condition", t["m_cond"])
    tables["Measure"] = a
    a = blank("HospitalType", len(t["types"]))
    put_str(a, "HospitalType", "desc", t["types"])
    tables["HospitalType"] = a
    a = blank("Hospital", H)
    put_fk(a, "Hospital", "loc", "Place", t["h_place"])
    put_fk(a, "Hospital", "type", "HospitalType", t["h_type"])
    put_str(a, "Hospital", "provider", t["provider"])
    put_str(a, "Hospital", "name", t["name"])
    put_str(a, "Hospital", "addr", t["addr"])
    put_str(a, "Hospital", "phone", t["phone"])
    put_str(a, "Hospital", "owner", [t["owners"][i] for i in t["h_owner"]])
    put_str(a, "Hospital", "zip", t["zip"])
    put_str(a, "Hospital", "service", [t["services"][i] for i in t["h_service"]])
    tables["Hospital"] = a
    ir.refresh()
    snap = {"tables": {}, "assignment": {}, "params": {}}
    for name, arr in tables.items():
        n = arr.shape[1]
        snap["tables"][name] = (np.arange(1, n + 1, dtype=np.int64), arr, 1.0, 0.0)
    rec = model.classes["Record"]
    snap["assignment"][rec.names["hosp"] - 1] = (t["row_h"] + 1).astype(np.int64)
    snap["assignment"][rec.names["metric"] - 1] = (t["row_m"] + 1).astype(np.int64)
    return snap


def build_synthetic_hospital(n_rows: int, seed: int = 20260924, **kw):
    """(model, query, dirty, truth, ir, obs, snapshot) for the synthetic hospital-schema table."""
    from .schemas import build_hospital
    dirty, truth = generate_hospital(n_rows, seed, **kw)
    # make sure every clean entity value is an option of its column (append if typos hid it)
    model, query = build_hospital(dirty)
    ds = M.ObservedDataset(query, dirty)
    ir = FlatIR(model, [ds])
    obs = ir.encode_observations(ds)
    snap = truth_snapshot(model, ir, truth)
    return model, query, dirty, truth, ir, obs, snap


# ------------------------------------------------------------------------------------------
# Synthetic rents-schema tables (BASELINE.json configs[4]): N rent
# observations over `n_counties` counties in 51 states and ~471 county keys; five AddTypos string
# columns (county name + four more name-like attributes, 10-35 characters, max_typos = 2), room
# type in 5, unit in 2 (1 % of the rents quoted in thousands), rent ~ N(mu[state, key, br], 150)
# with mu ~ N(1500, 1000) clipped at 300, 10 % missing room type, 10 % missing state, every string
# cell corrupted with probability `typo_rate` by one typo (a typo never touches the two characters
# the county key is made of: load_data.jl:9 derives the key from the observed name).
# ------------------------------------------------------------------------------------------
def generate_rents5(n_rows: int, seed: int = 20260925, n_counties: int = 3000, n_states: int = 51, typo_rate: float = 0.05):
    from .schemas.rents import EXTRA_ATTRS, EXTRA_COLUMNS, ROOM_TYPES, county_key
    rng = np.random.default_rng(seed)
    n_counties = max(4, min(n_counties, n_rows // 2))
    states = _rand_words(rng, n_states, 2, 2, 2, LETTERS.upper())
    firsts, lasts = LETTERS[:24], LETTERS[:20]            # 480 possible keys, nearly all of them used (the real data has 471)
    names, seen = [], set()
    while len(names) < n_counties:
        L = int(rng.integers(4, 12))
        w = firsts[rng.integers(0, len(firsts))] + "".join(LETTERS[c] for c in rng.integers(0, 26, size=L - 2)) + lasts[rng.integers(0, len(lasts))]
        tail = _rand_words(rng, 1, 12, 6, 22)[0]
        s = f"{w} {tail}"[:35]
        if len(s) >= 10 and s not in seen:
            seen.add(s); names.append(s)
    extras = {a: _rand_words(rng, n_counties, 20, 10, 35) for a in EXTRA_ATTRS}
    c_state = rng.integers(0, n_states, size=n_counties)
    keys = [county_key(s) for s in names]
    w = 1.0 / np.arange(1, n_counties + 1) ** 0.7
    row_c = rng.choice(n_counties, size=n_rows, p=w / w.sum())
    row_c[:n_counties] = rng.permutation(n_counties)      # every county is referenced, and by a clean row (below)
    row_br = rng.integers(0, len(ROOM_TYPES), size=n_rows)
    row_unit = (rng.random(n_rows) < 0.01).astype(np.int64)
    mu: Dict[Tuple[int, int], float] = {}
    cb = row_c.astype(np.int64) * len(ROOM_TYPES) + row_br
    ucb = np.unique(cb)
    mu_of = dict(zip(ucb.tolist(), np.maximum(300.0, rng.normal(1500.0, 1000.0, size=len(ucb))).tolist()))
    # one mean per (state, key, br): counties that share state and key share it
    for code in ucb.tolist():
        c, br = divmod(code, len(ROOM_TYPES))
        mu.setdefault((states[c_state[c]], keys[c], br), mu_of[code])
    mean = np.array([mu[(states[c_state[c]], keys[c], br)] for c, br in zip(row_c.tolist(), row_br.tolist())])
    rent = np.round(np.maximum(50.0, rng.normal(mean, 150.0)))
    obs_rent = np.where(row_unit == 1, rent / 1000.0, rent)

    def gather(values, idx):
        return np.array(values, dtype=object)[idx].tolist()

    clean = {"County": gather(names, row_c), "State": gather(states, c_state[row_c]), "Room Type": gather(ROOM_TYPES, row_br),
             "Monthly Rent": rent.tolist()}
    for a in EXTRA_ATTRS:
        clean[EXTRA_COLUMNS[a]] = gather(extras[a], row_c)
    dirty: Dict[str, List] = {"Monthly Rent": obs_rent.tolist()}
    for col in ["County"] + [EXTRA_COLUMNS[a] for a in EXTRA_ATTRS]:
        out = list(clean[col])
        hit = np.nonzero(rng.random(n_rows) < typo_rate)[0]
        hit = hit[hit >= n_counties]
        for i in hit.tolist():
            for _ in range(4):
                t = _typo(rng, out[i])
                if col != "County" or (len(t) >= 10 and county_key(t) == county_key(out[i]) and t.split()[0] == t.split()[0].strip()):
                    out[i] = t
                    break
        dirty[col] = out
    st = list(clean["State"]); br = list(clean["Room Type"])
    for i in np.nonzero(rng.random(n_rows) < 0.10)[0].tolist():
        if i >= n_counties:
            st[i] = None
    for i in np.nonzero(rng.random(n_rows) < 0.10)[0].tolist():
        if i >= n_counties:
            br[i] = None
    dirty["State"] = st; dirty["Room Type"] = br
    truth = dict(states=states, names=names, extras=extras, c_state=c_state, keys=keys, row_c=row_c, row_br=row_br, row_unit=row_unit,
                 mu=mu, clean=clean)
    return dirty, truth


def rents5_truth_snapshot(model: M.PCleanModel, ir: FlatIR, truth: dict) -> dict:
    from .schemas.rents import EXTRA_ATTRS, ROOM_TYPES, UNITS
    t = truth
    cm = model.classes["County"]
    n_normal = sum(1 for x in cm.nodes if not isinstance(x, M.ExternalLikelihoodNode))
    C = len(t["names"])
    a = np.zeros((n_normal, C), dtype=VALUE_DTYPE)
    a["tag"] = VAL_ABSENT

    def put(name, values):
        v = cm.names[name] - 1
        a[v]["tag"] = VAL_STR
        a[v]["i"] = np.array([ir.intern_string(s) for s in values], dtype=np.int32)

    put("countykey", t["keys"]); put("name", t["names"])
    for attr in EXTRA_ATTRS:
        put(attr, t["extras"][attr])
    put("state", [t["states"][i] for i in t["c_state"]])
    obs = model.classes["Obs"]
    snap = {"tables": {"County": (np.arange(1, C + 1, dtype=np.int64), a, 1.0, 0.0)}, "assignment": {}, "params": {}, "rowcells": {}}
    snap["assignment"][obs.names["county"] - 1] = (t["row_c"] + 1).astype(np.int64)
    n = len(t["row_c"])
    brc = np.zeros(n, dtype=VALUE_DTYPE); brc["tag"] = VAL_STR
    br_ids = np.array([ir.intern_string(s) for s in ROOM_TYPES], dtype=np.int32)
    brc["i"] = br_ids[t["row_br"]]
    snap["rowcells"][obs.names["br"] - 1] = brc
    uc = np.zeros(n, dtype=VALUE_DTYPE); uc["tag"] = 6                      # VAL_XFORM
    unit_ids = np.array([ir.encode(u)[1] for u in UNITS], dtype=np.int32)
    uc["i"] = unit_ids[t["row_unit"]]
    snap["rowcells"][obs.names["unit"] - 1] = uc
    spec = ir.param_spec[("Obs", obs.names["avg_rent"])]
    for (state, key, br), val in t["mu"].items():
        slot = ir.slot_id.get((spec, f"{state}_{key}_{ROOM_TYPES[br]}"))
        if slot is not None:
            snap["params"][slot] = [float(val)]
    ir.refresh()
    return snap


def build_synthetic_rents(n_rows: int, seed: int = 20260925, **kw):
    """(model, query, dirty, truth, ir, obs, snapshot) for the synthetic rents-schema table (5 AddTypos columns)."""
    from .schemas.rents import add_county_key, build_rents5
    dirty, truth = generate_rents5(n_rows, seed, **kw)
    add_county_key(dirty)
    truth["clean"]["CountyKey"] = list(dirty["CountyKey"])
    model, query = build_rents5(dirty)
    ds = M.ObservedDataset(query, dirty)
    ir = FlatIR(model, [ds])
    obs = ir.encode_observations(ds)
    snap = rents5_truth_snapshot(model, ir, truth)
    return model, query, dirty, truth, ir, obs, snap
