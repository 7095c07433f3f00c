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
    # the reference's experimental knowledge-sharing mode (BUILD_OPENMP: frirl_omp_run + merge_rb, frirl_agent.c), mountaincar,
    # 3 and 4 agents: per-agent episode sequences (steps, return, rules) + agent 0's final rule base  
    import re
    for env, W in (("mountaincar", 3), ("mountaincar", 4), ("acrobot", 3), ("cartpole", 3)):
        exe = os.path.join(HERE, "_ref", f"ref_{env}_omp_crtrig")
        with tempfile.TemporaryDirectory() as tmp:
            runs = []
            for _ in range(2):                       # thread scheduling must not matter: two runs, identical results
                r = subprocess.run([exe], cwd=tmp, check=True, capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS=str(W)))
                eps = {}
                for l in r.stdout.splitlines():
                    m = re.match(r"#(\d+) Episode: (\d+)\tSteps: (\d+)\tReward: (?:\x1b\[[0-9;]*m)?(-?[0-9.]+)(?:\x1b\[[0-9;]*m)?\tRules: (\d+)", l)
                    if m:
                        eps.setdefault(int(m.group(1)), []).append((int(m.group(3)), float(m.group(4)), int(m.group(5))))
                runs.append((eps, open(os.path.join(tmp, f"{env}.frirlrb.txt"), "rb").read()))
            if runs[0] != runs[1]:
                sys.exit(f"{env} omp W={W}: the reference run is not deterministic")
            eps, txt = runs[0]
            rows = np.array([(a, s_, r_, n_) for a in sorted(eps) for (s_, r_, n_) in eps[a]], dtype=np.float64)
            np.savez_compressed(os.path.join(out, f"{env}_omp{W}.npz"), episodes=rows)
            open(os.path.join(out, f"{env}_omp{W}.frirlrb.txt"), "wb").write(txt)
            print(f"knowledge sharing, {env}, {W} agents: {len(rows)} episodes, agent 0 ends with {len(txt.splitlines())} rules")
    # a problem that is none of the examples (3 state dimensions, 5 actions, skip_rules = 0): tests/dropin/custom_problem.c, our own
    # program on the reference's public API, built against the reference by oracle/Makefile
    # custom6 / custom8: the same program with 5 / 7 state dimensions (K = 6 / 8 = FIVE_MAX_NUM_OF_UNIVERSES antecedents)
    for tag in ("custom", "custom6", "custom8"):
        exe = os.path.join(HERE, "_ref", tag + "_ref")
        with tempfile.TemporaryDirectory() as tmp:
            r = subprocess.run([exe], cwd=tmp, check=True, capture_output=True, text=True)
            open(os.path.join(out, tag + "_problem.out"), "w").write(r.stdout)
            shutil.copy(os.path.join(tmp, "custom.frirlrb.txt"), os.path.join(out, tag + "_problem.frirlrb.txt"))
            print(f"{tag} problem: {len(r.stdout.splitlines())} output lines")


if __name__ == "__main__":
    main()
