"""Regenerates tests/golden/*.json from the reference's own test sources (run in the build container,
where /root/reference exists; the GPU box only reads the committed JSON).

rope_testrope.json : the two 64-value expected sin arrays of TestCorrectness.TestRope
                     (jlama-tests/src/test/java/com/github/tjake/jlama/model/TestCorrectness.java:93-115),
                     the only golden vectors the reference holds for the hot path.
"""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/jlama-tests/src/test/java/com/github/tjake/jlama/model/TestCorrectness.java"


def main():
    src = open(REF).read()
    blk = src[src.index("public void TestRope()"):src.index("public void testRope2()")]
    arrs = re.findall(r"new double\[\] \{(.*?)\};", blk, flags=re.S)
    vals = [[float(v) for v in a.replace("\n", " ").split(",") if v.strip()] for a in arrs]
    assert len(vals) == 2 and all(len(v) == 64 for v in vals)
    json.dump({"source": "jlama-tests/src/test/java/com/github/tjake/jlama/model/TestCorrectness.java:93-115 (TestRope)",
               "call": "VectorMath.precomputeFreqsCis(128, 8192, 10000.0, 1.0)", "tolerance": 1e-4,
               "sin_position_1": vals[0], "sin_position_64": vals[1]},
              open(os.path.join(HERE, "rope_testrope.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
