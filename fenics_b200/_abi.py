"""Enum values of the C ABI, parsed from include/fenics_b200.h (single source of truth)."""
import os
import re

HEADER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "fenics_b200.h")


def _parse(path):
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for body in re.findall(r"typedef\s+enum\s*\{(.*?)\}", src, flags=re.S):
        nxt = 0
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, val = [x.strip() for x in item.split("=")]
                nxt = int(val, 0)
            else:
                name = item
            out[name] = nxt
            nxt += 1
    return out


ENUM = _parse(HEADER)
globals().update(ENUM)

NSPACES = ENUM["FX_NSPACES"]
NFIELDS = ENUM["FX_NFIELDS"]
NPARAMS = ENUM["FX_NPARAMS"]
SPACE_ID = {"W": ENUM["FX_SPACE_W"], "Zc": ENUM["FX_SPACE_ZC"], "Q": ENUM["FX_SPACE_Q"],
            "Zd": ENUM["FX_SPACE_ZD"], "Qt": ENUM["FX_SPACE_QT"], "V": ENUM["FX_SPACE_V"], "Vs": ENUM["FX_SPACE_VS"]}
SPACE_LD = {"W": 15, "Zc": 18, "Q": 3, "Zd": 3, "Qt": 1, "V": 12, "Vs": 6}
