"""CPU: include/splade_b200.h is a plain C header and libsplade_b200.so is callable from C without Python or
torch - the boundary a cgo / JNI / ctypes stub on the reference side would bind.  A C99 program is compiled with
gcc against the header, linked with the shared library and run: ABI version, workspace sizing, and the host-side
emitters (no device calls here)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "opensearch-neural-pre-train_b200")

C_SOURCE = r'''
#include <stdio.h>
#include <string.h>
#include "splade_b200.h"

int main(void) {
  printf("abi=%d\n", splade_abi_version());
  printf("ws=%zu\n", splade_workspace_bytes(SPLADE_OP_HEAD_BWD, 64, 256, 768, 50000, 0));
  printf("win=%zu\n", splade_peer_window_bytes(64, 50000, 8));

  /* two documents, k = 2: compact top-k buffers as splade_topk leaves them on the host */
  const int32_t ids[4] = {1, 0, 2, -1};
  const float vals[4] = {1.5f, 0.25f, 0.125f, 0.0f};
  const int32_t counts[2] = {2, 1};
  const char tokens[] = "aabbcc";
  const int64_t tok_off[4] = {0, 2, 4, 6};
  const char doc_ids[] = "d0d1";
  const int64_t id_off[3] = {0, 2, 4};
  char out[512];
  int64_t offs[3];
  size_t need = 0;
  int rc = splade_emit_bulk(ids, vals, counts, 2, 2, tokens, tok_off, 3, -1, "idx", "sparse_embedding", doc_ids,
                            id_off, NULL, NULL, 0, out, sizeof(out), offs, &need);
  if (rc != SPLADE_OK) { printf("emit_bulk failed: %s\n", splade_last_error()); return 1; }
  fwrite(out, 1, need, stdout);

  /* error path: a too small buffer reports the size it needs and a message, and nothing aborts */
  rc = splade_emit_bulk(ids, vals, counts, 2, 2, tokens, tok_off, 3, -1, "idx", "sparse_embedding", doc_ids,
                        id_off, NULL, NULL, 0, out, 8, offs, &need);
  printf("small rc=%d need=%zu msg=%s\n", rc, need, strstr(splade_last_error(), "too small") ? "too small" : "?");
  /* null arguments are an error code, not a crash */
  rc = splade_head_fwd(NULL, NULL, NULL, NULL, 1, 1, 64, 8, NULL, NULL, NULL, NULL, 0, NULL);
  printf("null rc_nonzero=%d\n", rc != SPLADE_OK);
  return 0;
}
'''


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_header_is_c99_and_library_is_callable_from_c(tmp_path):
    lib = os.path.join(PKG, "libsplade_b200.so")
    assert os.path.exists(lib), "run __graft_entry__.build() first"
    src = tmp_path / "abi_smoke.c"
    src.write_text(C_SOURCE)
    exe = tmp_path / "abi_smoke"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe), "-L", PKG, "-l:libsplade_b200.so", f"-Wl,-rpath,{PKG}"],
                   check=True, capture_output=True, text=True)
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
    lines = run.stdout.splitlines()
    assert lines[0] == "abi=5"
    assert int(lines[1].split("=")[1]) >= 2 * 64 * 50000 * 4
    assert int(lines[2].split("=")[1]) == 512 + 2 * (2 * 64 * 50000 * 4 + 64 * 8 * 64 * 4 + 256)
    assert lines[3] == '{"index":{"_index":"idx","_id":"d0"}}'
    assert lines[4] == '{"sparse_embedding":{"bb":1.5,"aa":0.25}}'
    assert lines[5] == '{"index":{"_index":"idx","_id":"d1"}}'
    assert lines[6] == '{"sparse_embedding":{"cc":0.125}}'
    assert lines[7].startswith("small rc=4 need=") and lines[7].endswith("msg=too small")
    assert lines[8] == "null rc_nonzero=1"
