"""End-to-end runs of the applications on the toy datasets through the launcher (the reference's tests/run_apps.sh:
every app, 2 processes, a few iterations, success = clean exit + sane output). ComplEx, DSGD and word2vec are run by
the flag tests at the end of the file."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = os.path.join(ROOT, "data")

CASES = {
    "kge_rescal": (["-m", "adapm_b200.apps.kge", "--", "--dataset", os.path.join(D, "kge") + "/", "--num_entities", "280",
                    "--num_relations", "112", "--embed_dim", "4", "--num_epochs", "1", "--algorithm", "RESCAL"],
                   r"\[kge\] epoch 1: bce loss [\d.]+"),
    "mf_columnwise": (["-m", "adapm_b200.apps.mf", "--", "--dataset", os.path.join(D, "mf", "train.mmc"), "--rank", "2",
                       "--epochs", "2", "--algorithm", "columnwise"], r"\[mf\] epoch 1: local squared error [\d.]+"),
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


def _launch(args, world=2):
    r = subprocess.run([sys.executable, "-m", "adapm_b200.launch", "-s", str(world), "--backend", "cpu"] + args, cwd=ROOT,
                       capture_output=True, text=True, timeout=300)
    return r.returncode, r.stdout + r.stderr


def test_word2vec_reference_flags_and_checkpoint_roundtrip(tmp_path):
    base = ["-m", "adapm_b200.apps.word2vec", "--", "--input_file", os.path.join(D, "lm", "small.txt"), "--embed_dim", "16",
            "--num_iterations", "1", "--min_count", "5", "--batch_pairs", "8192", "--negative", "3"]
    ck = str(tmp_path / "vec")
    rc, out = _launch(base + ["--write_results", "1", "--output_file", ck, "--enforce_random_keys", "1", "--sync_push", "1"])
    assert rc == 0 and os.path.exists(ck + ".epoch.0"), out[-2000:]
    rc, out = _launch(base + ["--init_model", ck + ".epoch.0", "--enforce_full_replication", "1", "--data_words", "20000",
                              "--num_threads", "2", "--clustered_input", "0", "--debug_mode", "0"])
    assert rc == 0, out[-2000:]
    m = re.search(r"initialised (\d+) of (\d+) words", out)
    assert m and m.group(1) == m.group(2), out[-2000:]
    assert re.search(r"\[rank 0\] local pulls 100% ", out), out[-1500:]      # everything replicated everywhere


def test_kge_reference_flags(tmp_path):
    out_dir = str(tmp_path / "emb")
    rc, out = _launch(["-m", "adapm_b200.apps.kge", "--", "--dataset", os.path.join(D, "kge") + "/", "--num_entities", "280",
                       "--num_relations", "112", "--embed_dim", "8", "--num_epochs", "2", "--eval_initial", "1",
                       "--eval_truncate_va", "50", "--max_N", "400", "--init_parameters", "uniform{-0.1/0.1}",
                       "--enforce_random_keys", "1", "--write_embeddings", out_dir, "--async_push", "0"])
    assert rc == 0, out[-2000:]
    assert "[kge] initial valid: {'mrr_s'" in out and "'mrr':" in out and "'n': 50" in out, out[-2000:]
    assert os.path.exists(os.path.join(out_dir, "export.epoch.2.entities.bin"))
    rc, out = _launch(["-m", "adapm_b200.apps.kge", "--", "--dataset", os.path.join(D, "kge") + "/", "--num_entities", "280",
                       "--num_relations", "112", "--embed_dim", "8", "--num_epochs", "1", "--read_partitioned_dataset", "1",
                       "--num_threads", "2", "--init_parameters", "normal{0/0.05}"])
    assert rc == 0 and "[kge] epoch 1: bce loss" in out, out[-2000:]


def test_mf_reference_flags(tmp_path):
    pre = str(tmp_path / "init_")
    rc, out = _launch(["-m", "adapm_b200.apps.mf", "--", "--dataset", os.path.join(D, "mf", "train.mmc"), "--rank", "2",
                       "--epochs", "3", "--init_parameters", "1", "--write_generated_factors", pre, "--wor_blocks", "0",
                       "--increase_step_factor", "1.1", "--decrease_step_factor", "0.4", "--signal_intent_rows", "1"])
    assert rc == 0, out[-2000:]
    assert re.search(r"\[mf\] epoch 2: local squared error [\d.]+, eps [\d.e-]+, test rmse [\d.]+", out), out[-2000:]
    # the factors that were written are the ones that were read (data/mf/W.mma)
    want = open(os.path.join(D, "mf", "W.mma")).read().split()[-12:]
    got = open(pre + "W.mma").read().split()[-12:]
    assert [round(float(a), 4) for a in want] == [round(float(b), 4) for b in got]
    rc, out = _launch(["-m", "adapm_b200.apps.mf", "--", "--dataset", os.path.join(D, "mf", "train.mmc"), "--rank", "2",
                       "--epochs", "2", "--algorithm", "plain", "--enforce_random_keys", "1", "--early_stop", "10",
                       "--wor_points", "0", "--enforce_full_replication", "1", "--max_runtime", "100", "--compute_loss", "1"])
    assert rc == 0 and "[mf] epoch 1" in out, out[-2000:]


def test_kge_and_mf_with_double_values(tmp_path):
    """--value_type double = the reference's ValT for kge and mf (kge.cc:33, mf.cc): float64 rows through Pull / Push with
    the same update rule; the float32 and the float64 run of the same job agree to float32 rounding."""
    def loss_of(out, pattern):
        m = re.findall(pattern, out)
        assert m, out[-2000:]
        return float(m[-1])

    kge = ["-m", "adapm_b200.apps.kge", "--", "--dataset", os.path.join(D, "kge") + "/", "--num_entities", "280",
           "--num_relations", "112", "--embed_dim", "8", "--num_epochs", "2", "--write_embeddings", str(tmp_path / "d_"),
           "--sys.techniques", "replication_only"]
    rc, out64 = _launch(kge + ["--value_type", "double"], world=1)
    assert rc == 0 and "protocol errors 0" in out64, out64[-2000:]
    rc, out32 = _launch(kge, world=1)
    assert rc == 0, out32[-2000:]
    l64, l32 = (loss_of(o, r"\[kge\] epoch 2: bce loss ([\d.]+)") for o in (out64, out32))
    assert abs(l64 - l32) <= 2e-3 * abs(l64), (l64, l32)
    mf = ["-m", "adapm_b200.apps.mf", "--", "--dataset", os.path.join(D, "mf", "train.mmc"), "--rank", "2", "--epochs", "3",
          "--init_parameters", "1", "--algorithm", "plain", "--wor_points", "0"]
    rc, out64 = _launch(mf + ["--value_type", "double"], world=1)
    assert rc == 0 and "protocol errors 0" in out64, out64[-2000:]
    rc, out32 = _launch(mf, world=1)
    assert rc == 0, out32[-2000:]
    e64, e32 = (loss_of(o, r"\[mf\] epoch 2: local squared error ([\d.]+)") for o in (out64, out32))
    assert abs(e64 - e32) <= 2e-3 * abs(e64) + 1e-6, (e64, e32)
    # two ranks, double values: the whole protocol (relocation + replication) on float64 rows
    rc, out = _launch(kge + ["--value_type", "double"], world=2)
    assert rc == 0 and "[kge] epoch 2: bce loss" in out and "protocol errors 0" in out, out[-2000:]
