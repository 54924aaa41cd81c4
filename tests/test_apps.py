"""End-to-end runs of the applications on the toy datasets through the launcher (the reference's tests/run_apps.sh:
every app, 2 processes, a few iterations, success = clean exit + sane output)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = os.path.join(ROOT, "data")

CASES = {
    "kge_complex": (["-m", "adapm_b200.apps.kge", "--", "--dataset", os.path.join(D, "kge") + "/", "--num_entities", "280",
                     "--num_relations", "112", "--embed_dim", "8", "--num_epochs", "2", "--eval_freq", "2"],
                    r"\[kge\] epoch 2: bce loss [\d.]+.*\[kge\] test: \{'mrr'"),
    "kge_rescal": (["-m", "adapm_b200.apps.kge", "--", "--dataset", os.path.join(D, "kge") + "/", "--num_entities", "280",
                    "--num_relations", "112", "--embed_dim", "4", "--num_epochs", "1", "--algorithm", "RESCAL"],
                   r"\[kge\] epoch 1: bce loss [\d.]+"),
    "mf_dsgd": (["-m", "adapm_b200.apps.mf", "--", "--dataset", os.path.join(D, "mf", "train.mmc"), "--rank", "2",
                 "--epochs", "3"], r"\[mf\] epoch 2: local squared error [\d.]+"),
    "mf_columnwise": (["-m", "adapm_b200.apps.mf", "--", "--dataset", os.path.join(D, "mf", "train.mmc"), "--rank", "2",
                       "--epochs", "2", "--algorithm", "columnwise"], r"\[mf\] epoch 1: local squared error [\d.]+"),
    "word2vec": (["-m", "adapm_b200.apps.word2vec", "--", "--input_file", os.path.join(D, "lm", "small.txt"), "--embed_dim",
                  "16", "--num_iterations", "1", "--min_count", "5", "--batch_pairs", "8192", "--negative", "3"],
                 r"\[w2v\] epoch 0 done: \d+ pairs"),
    "ctr": (["-m", "adapm_b200.apps.ctr", "--", "--steps", "3", "--batch_size", "256", "--num_features", "5000",
             "--num_fields", "6", "--precision", "fp32"],
            r"\[ctr\] step 3: loss [\d.]+"),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_app_runs(name):
    args, pattern = CASES[name]
    r = subprocess.run([sys.executable, "-m", "adapm_b200.launch", "-s", "2", "--backend", "cpu"] + args, cwd=ROOT,
                       capture_output=True, text=True, timeout=300)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert re.search(pattern, out, re.S), out[-3000:]
    assert "protocol errors 0" in out, out[-1500:]
