"""Generator of the reference's interface-test configurations (the loop of tests/Makefile:146-189 of the
reference, `make tests_primme_interface`): every preset method x Laplacian size x number of pairs x target x
extraction, each checked by the reference's driver against the STORED solution
tests/sol_testi-<n>-<nevals>-<target>_double.  The configurations are text files for the reference's own
config reader; they are generated into a scratch directory at test time, not committed."""
import os

METHODS = ["DEFAULT_METHOD", "DYNAMIC", "DEFAULT_MIN_TIME", "DEFAULT_MIN_MATVECS", "Arnoldi", "GD_plusK", "GD_Olsen_plusK",
           "JD_Olsen_plusK", "JDQR", "JDQMR", "JDQMR_ETol", "STEEPEST_DESCENT", "LOBPCG_OrthoBasis", "LOBPCG_OrthoBasis_Window"]
SIZES = [0, 1, 2, 3, 4, 5, 6, 7, 10, 100]
NEVALS = [0, 1, 2, 3, 4, 5, 6, 15, 100]
TARGETS = ["primme_smallest", "primme_largest", "primme_closest_abs", "primme_closest_geq"]
PROJS = ["primme_proj_RR", "primme_proj_refined"]


def laplace_mtx(n):
    lines = ["%%MatrixMarket matrix coordinate real symmetric", f"{n} {n} {2 * n - 1 if n > 0 else 0}"]
    for i in range(1, n + 1):
        lines.append(f"{i} {i} 2.0")
        if i != n:
            lines.append(f"{i} {i + 1} -1.0")
    return "\n".join(lines) + "\n"


def skipped(target, n, nevals, method, proj):
    """the exclusions of the reference's Makefile"""
    if target == "primme_closest_geq" and (n, nevals) in ((4, 4), (5, 5), (6, 6), (7, 7)):
        return True
    if target in ("primme_closest_geq", "primme_closest_leq") and n == 100 and method.startswith("LOBPCG"):
        return True
    if target.startswith("primme_closest"):
        if proj == "primme_proj_RR" and method.startswith("LOBPCG"):
            return True
        if method.startswith("STEEPEST_DESCENT") or method.startswith("Arnoldi") or method.startswith("GD"):
            return True
        return False
    return proj != "primme_proj_RR"


def configs():
    for method in METHODS:
        for n in SIZES:
            for nevals in NEVALS:
                if nevals > n:
                    continue
                for target in TARGETS:
                    for proj in PROJS:
                        if skipped(target, n, nevals, method, proj):
                            continue
                        name = f"testi-{n}-{method}-{nevals}-{target}-{proj}"
                        text = "\n".join([f"driver.matrixFile = laplace{n}.mtx",
                                          f"driver.checkXFile = tests/sol_testi-{n}-{nevals}-{target}_double",
                                          "driver.PrecChoice = noprecond", f"primme.numEvals = {nevals}", "primme.eps = 1e-6",
                                          "primme.numTargetShifts = 1", "primme.targetShifts  = 0.5", f"primme.target = {target}",
                                          f"primme.projection.projection = {proj}", "primme.maxMatvecs = 50000",
                                          f"method = PRIMME_{method}"]) + "\n"
                        yield name, method, text


def write_all(directory):
    os.makedirs(directory, exist_ok=True)
    for n in SIZES:
        open(os.path.join(directory, f"laplace{n}.mtx"), "w").write(laplace_mtx(n))
    names = []
    for name, method, text in configs():
        open(os.path.join(directory, name + ".F"), "w").write(text)
        names.append((name, method))
    return names
