#!/usr/bin/env python3
"""Extract the reference's own known-answer vectors into tests/golden/reference_kats.json.

Runs in the build container only (needs /root/reference).  Sources:
  * edge-order CRC-32s            src/codes/mod.rs:521-523
  * encode parity blocks          src/encoder.rs:363-526  (data bytes 0,1,2,...)
  * converter vector              src/decoder.rs:553-605
  * doc-test vectors              src/lib.rs:135-143
  * literal length table          src/codes/mod.rs:109-241
  * header size macros            capi/include/labrador_ldpc.h:45-113
"""
import json
import os
import re

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODES = ["TC128", "TC256", "TC512", "TM1280", "TM1536", "TM2048", "TM5120", "TM6144", "TM8192"]


def main():
    out = {"source": "adamgreig/labrador-ldpc v1.2.1 test vectors", "codes": CODES}

    mod = open(os.path.join(REF, "src/codes/mod.rs")).read()
    m = re.search(r"let crc_results = \[(.*?)\];", mod, re.S)
    out["edge_crc32"] = [int(x, 16) for x in re.findall(r"0x([0-9A-Fa-f]+)", m.group(1))]
    assert len(out["edge_crc32"]) == 9

    # literal CodeParams table
    params = {}
    for code in CODES:
        m = re.search(r"pub const %s_PARAMS: CodeParams = CodeParams \{(.*?)\};" % code, mod, re.S)
        fields = {}
        for name, expr in re.findall(r"(\w+):\s*([^,]+),", m.group(1)):
            fields[name] = int(eval(expr.replace("/", "//")))
        params[code] = fields
    out["params"] = params

    enc = open(os.path.join(REF, "src/encoder.rs")).read()
    kats = {}
    for code in CODES:
        m = re.search(r"test_encode!\(LDPCCode::%s,\s*\[(.*?)\]\);" % code, enc, re.S)
        kats[code] = [int(x, 16) for x in re.findall(r"0x([0-9A-Fa-f]{2})", m.group(1))]
        assert len(kats[code]) == (params[code]["n"] - params[code]["k"]) // 8
    out["encode_parity"] = kats

    dec = open(os.path.join(REF, "src/decoder.rs")).read()
    m = re.search(r"fn test_hard_to_llrs\(\).*?let hard = vec!\[(.*?)\];.*?assert_eq!\(llrs, vec!\[(.*?)\]\);", dec, re.S)
    out["convert_hard"] = [int(x) for x in re.findall(r"\d+", m.group(1))]
    signs = re.findall(r"(-?)llr", m.group(2))
    # llr = -1.0; "llr" -> -1.0, "-llr" -> +1.0
    out["convert_llrs"] = [1.0 if s == "-" else -1.0 for s in signs]
    assert len(out["convert_hard"]) == 16 and len(out["convert_llrs"]) == 128

    out["doctest_tc128_codeword"] = [0, 1, 2, 3, 4, 5, 6, 7, 0x34, 0x99, 0x98, 0x87, 0x94, 0xE1, 0x62, 0x56]
    lib = open(os.path.join(REF, "src/lib.rs")).read()
    assert "0x34, 0x99, 0x98, 0x87, 0x94, 0xE1, 0x62, 0x56" in lib
    assert "0x5662E19487989934" in lib
    out["doctest_tc128_u64_parity_word"] = 0x5662E19487989934

    hdr = open(os.path.join(REF, "capi/include/labrador_ldpc.h")).read()
    macros = {}
    for name, val in re.findall(r"#define (LABRADOR_LDPC_[A-Z0-9_]+_T[CM]\d+)\s+\((\d+)\)", hdr):
        macros[name] = int(val)
    out["header_macros"] = macros

    path = os.path.join(ROOT, "tests", "golden", "reference_kats.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
