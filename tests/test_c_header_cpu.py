"""include/yama_b200.h must be plain C (it is what a C host like multiz includes) and every host-only entry point must be
callable from C without a GPU.  Compiles a small C program with gcc against the library and runs it."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROG = r'''
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "yama_b200.h"
int main(void) {
    int32_t LB[4] = {0, 0, 0, 0}, RB[4] = {12, 12, 12, 12};
    char msg[128];
    if (yb_check_band(3, 12, LB, RB, msg, sizeof msg) != 52) return 1;
    RB[1] = 5;
    if (yb_check_band(3, 12, LB, RB, msg, sizeof msg) != YB_ERR_BAND || strcmp(msg, "RB[1] - LB[1] < 10, 5 0 12")) return 2;
    int64_t cells[5] = {10, 10, 10, 10, 10}, cuts[3];
    if (yb_plan_split(5, cells, 2, cuts) != YB_OK || cuts[0] != 0 || cuts[2] != 5) return 3;
    /* column assembly from a packed script: ops (reference order, reversed) C, D, I  ->  I then D then C forward */
    uint8_t A[2] = {'A', 'C'}, B[2] = {'G', 'T'}, script[1] = {(uint8_t)(0 | (2 << 2) | (1 << 4))}, out[6], ops[3];
    yb_job job = {1, 2, 1, 2, A, B, LB, RB};
    yb_result res;
    memset(&res, 0, sizeof res);
    res.m_new = 3; res.script = script;
    if (yb_assemble(&job, &res, out) != YB_OK || memcmp(out, "-GA-CT", 6)) return 4;
    if (yb_script_unpack(&res, ops) != YB_OK || ops[0] != 0 || ops[1] != 2 || ops[2] != 1) return 5;
    yb_ctx *ctx = NULL;
    int rc = yb_create(NULL, 0, &ctx);          /* fails cleanly without a GPU, works with one */
    if (rc == YB_OK) yb_destroy(ctx); else if (rc != YB_ERR_CUDA || ctx != NULL) return 6;
    puts("ok");
    return 0;
}
'''


def test_header_is_c_and_host_entry_points_work(tmp_path):
    src = tmp_path / "abi.c"
    src.write_text(PROG)
    exe = tmp_path / "abi"
    libdir = os.path.join(ROOT, "multiz_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", str(exe), "-L", libdir, "-lyama_b200", "-Wl,-rpath," + libdir], check=True)
    p = subprocess.run([str(exe)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0 and p.stdout.strip() == b"ok", (p.returncode, p.stderr)
