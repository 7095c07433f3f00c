#!/usr/bin/env python3
"""oracle/make_golden.py -- TEST INFRASTRUCTURE: regenerate tests/golden/.

Runs the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile from
/root/reference, linked with the correctly-rounded sin/cos shim) and commits
its outputs as fixtures that travel to the GPU box (where /root/reference does
not exist):

  tests/golden/<env>.frirlrb.txt   final rule base of the construct run.  Checked
                                   here to be BYTE-IDENTICAL to the reference's
                                   own tests/orig/frirl_example_<env>.frirlrb.txt
                                   (tests/_test.sh:8-17 diffs exactly these).
  tests/golden/<env>.frirlrb.bin   the binary dump (frirl_utils.c:151-203).
  tests/golden/mountaincar_reduced.frirlrb.txt, mountaincar_reduction.log
                                   the shipped mountaincar run: reduction of the golden rule base.
  tests/golden/<env>_calls.npz     per-call known-answer vectors recorded by
                                   oracle/ref_harness.c (record stream, see there)
                                   + the md5 of the text golden.

Usage:  make -C oracle ref && python oracle/make_golden.py
"""
import hashlib, os, shutil, subprocess, sys, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("REF", "/root/reference")
ENVS = ["mountaincar", "acrobot", "cartpole"]


def main():
    out = os.path.join(REPO, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for env in ENVS:
        exe = os.path.join(HERE, "_ref", f"ref_{env}_crtrig")
        with tempfile.TemporaryDirectory() as tmp:
            rec = os.path.join(tmp, "rec.bin")
            subprocess.run([exe], cwd=tmp, check=True,
                           env=dict(os.environ, REF_MODE="record", REF_OUT=rec))
            txt = open(os.path.join(tmp, f"{env}.frirlrb.txt"), "rb").read()
            orig = open(os.path.join(REF, "tests", "orig", f"frirl_example_{env}.frirlrb.txt"), "rb").read()
            if txt != orig:
                sys.exit(f"{env}: reference run is NOT byte-identical to tests/orig")
            shutil.copy(os.path.join(tmp, f"{env}.frirlrb.txt"), os.path.join(out, f"{env}.frirlrb.txt"))
            shutil.copy(os.path.join(tmp, f"{env}.frirlrb.bin"), os.path.join(out, f"{env}.frirlrb.bin"))
            stream = np.fromfile(rec, dtype="<f8")
            np.savez_compressed(os.path.join(out, f"{env}_calls.npz"), stream=stream,
                                md5=np.array(hashlib.md5(txt).hexdigest()))
            print(f"{env}: byte-identical to tests/orig (md5 {hashlib.md5(txt).hexdigest()}), "
                  f"{stream.size} doubles recorded")
    # mountaincar exactly as shipped (construct_rb = 0, reduce_rb = 1, mountaincar.c:240-242): rule-base REDUCTION
    # (frirl_sequential_run.c:170-350) of the constructed golden rule base, 110 -> 28 rules
    exe = os.path.join(HERE, "_ref", "ref_mountaincar_crtrig")
    with tempfile.TemporaryDirectory() as tmp:
        r = subprocess.run([exe], cwd=tmp, check=True, capture_output=True, text=True,
                           env=dict(os.environ, REF_MODE="shipped", REF_VERBOSE="1",
                                    REF_RBFILE=os.path.join(out, "mountaincar.frirlrb.bin")))
        shutil.copy(os.path.join(tmp, "mountaincar.frirlrb.txt"), os.path.join(out, "mountaincar_reduced.frirlrb.txt"))
        keep = [l for l in r.stdout.splitlines() if l.startswith(("Reduction Episode", "Smallest", "Loaded"))]
        keep = [l.split(" rules from")[0] if l.startswith("Loaded") else l for l in keep]
        open(os.path.join(out, "mountaincar_reduction.log"), "w").write("\n".join(keep) + "\n")
        print(f"mountaincar reduction: {len(open(os.path.join(out, 'mountaincar_reduced.frirlrb.txt')).readlines())} rules, {len(keep)} log lines")
    # a problem that is none of the examples (3 state dimensions, 5 actions, skip_rules = 0): tests/dropin/custom_problem.c, our own
    # program on the reference's public API, built against the reference by oracle/Makefile
    exe = os.path.join(HERE, "_ref", "custom_ref")
    with tempfile.TemporaryDirectory() as tmp:
        r = subprocess.run([exe], cwd=tmp, check=True, capture_output=True, text=True)
        open(os.path.join(out, "custom_problem.out"), "w").write(r.stdout)
        shutil.copy(os.path.join(tmp, "custom.frirlrb.txt"), os.path.join(out, "custom_problem.frirlrb.txt"))
        print(f"custom problem: {len(r.stdout.splitlines())} output lines")


if __name__ == "__main__":
    main()
