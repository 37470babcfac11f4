"""Extracts the 66 expected Legendre-cache unique identifiers of the reference's own test
(/root/reference/src/tests/trans/test_trans_localcache.cc:264-360, CASE "ATLAS-256: Legendre coefficient expected unique
identifiers") together with the loop structure that pairs each string with a (domain, truncation, grid) case, and writes
tests/golden/legendre_cache_uids.json.  Run in the build container (the reference tree does not travel to the GPU box)."""
import json
import os
import re

SRC = "/root/reference/src/tests/trans/test_trans_localcache.cc"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    text = open(SRC).read()
    body = text[text.index('CASE("ATLAS-256: Legendre coefficient expected unique identifiers")'):]
    uids = re.findall(r'"(local-T\d+-[^"]+)"', body)
    assert len(uids) == 66, len(uids)
    grids = re.search(r"std::vector<std::string>\{([^}]*)\}", body).group(1)
    grids = re.findall(r'"([A-Z]\d+)"', grids)
    T = [int(x) for x in re.search(r"std::vector<int>\{([^}]*)\}", body).group(1).split(",")]
    cases = []
    it = iter(uids)
    for domain in ("global", "rectangular lon [-10, 10] lat [-20, 20]"):
        for t in T:
            for g in grids:
                cases.append({"domain": domain, "truncation": t, "grid": g, "uid": next(it)})
    json.dump({"source": "src/tests/trans/test_trans_localcache.cc:264-360", "flt": False, "cases": cases},
              open(os.path.join(HERE, "legendre_cache_uids.json"), "w"), indent=1)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
