"""CPU: include/s3g_b200.h is a plain C header and libs3g_b200.so links from C - the boundary the reference's
C++ glue (or cgo / JNI / any FFI) would bind.  A small C program is compiled with gcc -std=c99 -pedantic,
linked against the library and run: it only exercises entry points that need no GPU (sizes, argument checks)."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

C_SRC = r'''
#include <stdio.h>
#include <string.h>
#include "s3g_b200.h"

static char* no_alloc(void* user, size_t bytes) { (void)user; (void)bytes; return NULL; }

int main(void) {
    s3g_adam_tensor t;
    s3g_plane_desc pd;
    s3g_row_tensor rt;
    s3g_deform_net net;
    int rc;
    memset(&t, 0, sizeof t); memset(&pd, 0, sizeof pd); memset(&rt, 0, sizeof rt); memset(&net, 0, sizeof net);
    if (s3g_abi_version() != S3G_ABI_VERSION) return 10;
    if (strcmp(s3g_build_arch(), "sm_100a") != 0) return 11;
    if (s3g_geom_bytes(1000) == 0 || s3g_binning_bytes(1000) == 0 || s3g_image_bytes(64, 64) == 0) return 12;
    if (s3g_sort_temp_bytes(4096) == 0 || s3g_knn_workspace_bytes(1000) == 0) return 13;
    if (s3g_image_loss_workspace_bytes(1, 3, 64, 64) == 0) return 14;
    /* argument checks return S3G_ERR_ARG and leave a message; nothing is launched */
    rc = s3g_adam_step(1, NULL, 0.9, 0.999, 1e-15, NULL);
    if (rc != S3G_ERR_ARG || strlen(s3g_last_error()) == 0) return 20;
    t.numel = 4; t.step = 1;                                  /* null pointers */
    if (s3g_adam_step(1, &t, 0.9, 0.999, 1e-15, NULL) != S3G_ERR_ARG) return 21;
    if (s3g_plane_reg_forward(1, &pd, NULL, NULL, NULL) != S3G_ERR_ARG) return 22;
    if (s3g_gather_rows(1, &rt, 4, 5, NULL, NULL) != S3G_ERR_ARG) return 23;
    if (s3g_peer_reduce_scatter(1, 0, NULL, 8, NULL) != S3G_ERR_ARG) return 24;
    if (s3g_knn_mean_dist2(-1, NULL, NULL, NULL, NULL) != S3G_ERR_ARG) return 25;
    if (s3g_densify_stats(-1, NULL, NULL, NULL, NULL, NULL, NULL) != S3G_ERR_ARG) return 26;
    if (s3g_mark_visible(0, NULL, NULL, NULL, NULL, NULL) < 0) return 27;       /* P == 0 short-circuits */
    if (s3g_deform_forward_workspace_bytes(&net) != 0 && s3g_deform_forward_workspace_bytes(NULL) != 0) return 28;
    (void)no_alloc;
    printf("c abi ok\n");
    return 0;
}
'''


def test_header_is_plain_c_and_library_links_from_c(built_lib, tmp_path):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    from s3gaussian_b200 import _lib
    libdir = os.path.dirname(_lib.LIB_PATH)
    src = tmp_path / "abi.c"
    src.write_text(C_SRC)
    exe = tmp_path / "abi"
    cmd = [gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
           "-L", libdir, "-ls3g_b200", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    env = dict(os.environ)
    cuda_lib = "/usr/local/cuda/lib64"
    env["LD_LIBRARY_PATH"] = cuda_lib + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    r = subprocess.run([str(exe)], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "c abi ok" in r.stdout, (r.returncode, r.stdout, r.stderr)
