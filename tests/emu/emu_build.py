"""Builds the host emulation of sd_train_kernel (libsd_train_emu.so) and the torch-free GPU checker
(sd_train_check) into tests/emu/_build/.  Test infrastructure only."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build")
EMU_LIB = os.path.join(OUT, "libsd_train_emu.so")
CHECK_BIN = os.path.join(OUT, "sd_train_check")
# fmaf must be ONE rounding (the GPU's FFMA) and nothing else may be contracted
CXXFLAGS = ["-O2", "-ffp-contract=off", "-mfma", "-std=c++17", "-pthread"]


def _newer(target, deps):
    return (not os.path.exists(target)) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps)


def build_emu():
    os.makedirs(OUT, exist_ok=True)
    deps = [os.path.join(HERE, "sd_train_emu.cpp"), os.path.join(HERE, "cta_emu.h"),
            os.path.join(ROOT, "scopa_b200", "csrc", "ms_sd_train.cuh"),
            os.path.join(ROOT, "scopa_b200", "csrc", "ms_sd_avgpol.cuh"),
            os.path.join(ROOT, "scopa_b200", "csrc", "ms_sd_train_cluster.cuh"),
            os.path.join(ROOT, "scopa_b200", "csrc", "ms_sd_sample.cuh"),
            os.path.join(ROOT, "scopa_b200", "csrc", "ms_div.cuh")]
    if _newer(EMU_LIB, deps):
        subprocess.run(["g++"] + CXXFLAGS + ["-fPIC", "-shared", "-o", EMU_LIB, deps[0]], check=True)
    return EMU_LIB


STATE_LIB = os.path.join(OUT, "libms_state_host.so")
HOST_H = os.path.join(HERE, "host_intrinsics.h")   # intrinsic shims shared by the ms_*_host.cpp builds


def _cuda_root():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return os.path.dirname(os.path.dirname(os.path.realpath(nvcc)))


def build_state_host():
    """the product's rule header (csrc/ms_state.cuh) compiled for the host; needs only the CUDA headers (vector types)"""
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(HERE, "ms_state_host.cpp")
    deps = [src, HOST_H, os.path.join(ROOT, "scopa_b200", "csrc", "ms_state.cuh")]
    if _newer(STATE_LIB, deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", f"-I{_cuda_root()}/include", f"-I{HERE}", "-o", STATE_LIB, src],
                       check=True)
    return STATE_LIB


TEAM_LIB = os.path.join(OUT, "libms_team_host.so")


def build_team_host():
    """the product's 2v2 team code (csrc/ms_team.cu: rules AND the three kernels) compiled for the host"""
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(HERE, "ms_team_host.cpp")
    csrc = os.path.join(ROOT, "scopa_b200", "csrc")
    deps = [src, HOST_H] + [os.path.join(csrc, f) for f in ("ms_team.cu", "ms_state.cuh", "ms_common.cuh")]
    if _newer(TEAM_LIB, deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", f"-I{_cuda_root()}/include", f"-I{HERE}", "-o", TEAM_LIB, src],
                       check=True)
    return TEAM_LIB


FULL_LIB = os.path.join(OUT, "libms_full_host.so")


def build_full_host():
    """the product's 40-card Scopa code (csrc/ms_full.cu: rules AND the init/step/legal/evaluate/rollout kernels) for the host"""
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(HERE, "ms_full_host.cpp")
    csrc = os.path.join(ROOT, "scopa_b200", "csrc")
    deps = [src, HOST_H] + [os.path.join(csrc, f) for f in ("ms_full.cu", "ms_common.cuh")]
    if _newer(FULL_LIB, deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", f"-I{_cuda_root()}/include", f"-I{HERE}", "-o", FULL_LIB, src],
                       check=True)
    return FULL_LIB


ENV_LIB = os.path.join(OUT, "libms_env_host.so")


def build_env_host():
    """the product's env kernels (csrc/ms_env.cu: deal / full deck / step / legal / capture / keys / rollout) for the host"""
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(HERE, "ms_env_host.cpp")
    csrc = os.path.join(ROOT, "scopa_b200", "csrc")
    deps = [src, HOST_H] + [os.path.join(csrc, f) for f in ("ms_env.cu", "ms_state.cuh", "ms_common.cuh")]
    if _newer(ENV_LIB, deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", f"-I{_cuda_root()}/include", f"-I{HERE}", "-o", ENV_LIB, src],
                       check=True)
    return ENV_LIB


SOLVER_LIB = os.path.join(OUT, "libms_solver_host.so")


def build_solver_host():
    """the product's tree enumeration + vanilla-CFR kernels (csrc/ms_solver.cu) on the CTA emulator; IEEE double, no contraction"""
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(HERE, "ms_solver_host.cpp")
    csrc = os.path.join(ROOT, "scopa_b200", "csrc")
    deps = [src, HOST_H, os.path.join(HERE, "cta_emu.h"), os.path.join(HERE, "cta_emu_warp.h")] + [os.path.join(csrc, f) for f in ("ms_solver.cu", "ms_tree_walk.cuh",
                                                                                     "ms_static_walk.cuh", "ms_state.cuh", "ms_common.cuh", "ms_div.cuh")]
    if _newer(SOLVER_LIB, deps):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-pthread", "-fPIC", "-shared", "-w",
                        f"-I{_cuda_root()}/include", f"-I{HERE}", "-o", SOLVER_LIB, src], check=True)
    return SOLVER_LIB


MD_LIB = os.path.join(OUT, "libms_multideal_host.so")


def build_multideal_host():
    """the product's multi-deal MCCFR kernels (csrc/ms_multideal.cu) on the CTA emulator, table in host memory"""
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(HERE, "ms_multideal_host.cpp")
    csrc = os.path.join(ROOT, "scopa_b200", "csrc")
    deps = [src, HOST_H, os.path.join(HERE, "cta_emu.h"), os.path.join(HERE, "cta_emu_warp.h")] + [os.path.join(csrc, f) for f in ("ms_multideal.cu", "ms_tree_walk.cuh",
                                                                                     "ms_static_walk.cuh", "ms_state.cuh", "ms_common.cuh", "ms_div.cuh")]
    if _newer(MD_LIB, deps):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-pthread", "-fPIC", "-shared", "-w",
                        f"-I{_cuda_root()}/include", f"-I{HERE}", "-o", MD_LIB, src], check=True)
    return MD_LIB


SD_LIB = os.path.join(OUT, "libms_sdcfr_host.so")


def build_sdcfr_host():
    """the product's SDCFR kernels, fp32 path (csrc/ms_sdcfr.cu) on the CTA emulator; separate mul / add like --fmad=false"""
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(HERE, "ms_sdcfr_host.cpp")
    csrc = os.path.join(ROOT, "scopa_b200", "csrc")
    deps = [src, HOST_H, os.path.join(HERE, "cta_emu.h"), os.path.join(HERE, "cta_emu_warp.h")] + [os.path.join(csrc, f) for f in ("ms_sdcfr.cu", "ms_state.cuh", "ms_common.cuh", "ms_div.cuh")]
    if _newer(SD_LIB, deps):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-pthread", "-fPIC", "-shared", "-w",
                        f"-I{_cuda_root()}/include", f"-I{HERE}", "-o", SD_LIB, src], check=True)
    return SD_LIB


def build_check():
    """needs libscopa_b200.so (scopa_b200/_build.py) and the CUDA runtime headers; links both libraries by rpath"""
    build_emu()
    lib_dir = os.path.join(ROOT, "scopa_b200")
    src = os.path.join(HERE, "sd_train_check.cpp")
    deps = [src, EMU_LIB, os.path.join(lib_dir, "libscopa_b200.so")]
    if _newer(CHECK_BIN, deps):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        cuda = os.path.dirname(os.path.dirname(os.path.realpath(nvcc)))
        subprocess.run(["g++", "-O2", "-std=c++17", "-o", CHECK_BIN, src, f"-I{cuda}/include", f"-L{cuda}/lib64",
                        f"-L{lib_dir}", f"-L{OUT}", "-lscopa_b200", "-lsd_train_emu", "-lcudart", "-lm",
                        "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,$ORIGIN/../../../scopa_b200", f"-Wl,-rpath,{cuda}/lib64"],
                       check=True)
    return CHECK_BIN


if __name__ == "__main__":
    print(build_emu())
    print(build_state_host())
    print(build_team_host())
    print(build_full_host())
    print(build_env_host())
    print(build_solver_host())
    print(build_multideal_host())
    print(build_sdcfr_host())
    print(build_check())
