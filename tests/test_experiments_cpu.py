"""CPU smoke tests of the experiment drivers (argparse CLIs mirroring /root/reference/experiments/*)."""
import pickle

import torch


def test_baseline_throughput_cli_runs_on_cpu():
    from lah_b200.experiments.throughput import baseline_throughput as bt
    args = bt.make_parser().parse_args(["--block-type", "ffn", "--gpus", "-1", "--layers-per-gpu", "2", "--hid-dim", "16",
                                        "--batch-size", "4", "--batches-for-latency", "1", "--batches-for-throughput", "2",
                                        "--throughput-runs", "1", "--linspace-points", "2", "--max-ping", "0.0005"])
    rows = bt.run(args, printer=lambda *_: None)
    assert [r[0] for r in rows] == ["fast", "slow", "slow"] and all(r[3] > 0 for r in rows)


def test_convergence_runner_emulator_and_history_layout(tmp_path):
    """script form of the notebooks: async trainers with stale gradients; metrics pickle has the notebooks' layout"""
    from lah_b200.experiments.convergence import run as conv
    for setup in ("dmoe", "faulty", "largeffn"):
        args = conv.make_parser().parse_args(["--setup", setup, "--num-experts", "4", "--num-active", "2", "--num-trainers", "2",
                                              "--device", "cpu", "--total-steps", "6", "--eval-interval", "3", "--layer-dim", "16",
                                              "--delay-ms", "1", "--eval-batches", "1", "--eval-batch-size", "16",
                                              "--logdir", str(tmp_path)])
        result = conv.run(args, printer=lambda *_: None)
        assert len(result["train_history"]) >= 6 and result["val_history"]
        assert {"loss", "delay_steps"} <= set(result["train_history"][0])
        assert {"loss", "acc", "num_updates"} <= set(result["val_history"][-1])
    files = sorted(p.name for p in tmp_path.iterdir())
    assert "delay1ms_dmoe2outof4experts_seed1337.pkl" in files and "delay1ms_largeffn_seed1337.pkl" in files
    hist = pickle.load(open(tmp_path / "delay1ms_dmoe2outof4experts_seed1337.pkl", "rb"))
    assert set(hist) == {"train_history", "val_history"}


def test_throughput_cli_flag_parity():
    """same flags as the reference scripts"""
    from lah_b200.experiments.throughput import throughput_server, throughput_client, inbox_chain
    s = throughput_server.make_parser().parse_args(["-p", "1", "--gpu", "0", "--block-type", "ffn"])
    assert (s.handler_processes, s.hid_dim, s.max_batch_size, s.layers_per_gpu) == (256, 1024, 2048, 56)
    c = throughput_client.make_parser().parse_args(["-j", "64", "--hosts", "a:1", "--block-type", "transformer"])
    assert (c.batches_for_latency, c.batches_for_throughput, c.throughput_runs, c.batch_size, c.linspace_points,
            c.max_ping) == (10, 100, 10, 2048, 10, 0.2)
    i = inbox_chain.make_parser().parse_args([])
    assert (i.layers_per_gpu, i.jobs, i.block_type) == (56, 64, "transformer")


def test_inbox_chain_schedule_invariants():
    """pipeline schedule of the in-box throughput experiment: every rank always has exactly one wave, a wave walks the
    chain in order, and layer i is always executed by rank i % world"""
    from lah_b200.experiments.throughput.inbox_chain import chain_schedule
    for world, per_gpu in ((1, 5), (2, 3), (8, 56)):
        L = world * per_gpu
        position = {}
        for t in range(world, world + 2 * L):          # steady state (every wave has entered the chain)
            waves = set()
            for r in range(world):
                wave, layer, local = chain_schedule(t, r, world, L)
                assert layer % world == r and local == layer // world and 0 <= local < per_gpu
                waves.add(wave)
                if wave in position:
                    assert layer == (position[wave] + 1) % L      # the wave advanced by exactly one layer
                position[wave] = layer
            assert waves == set(range(world))               # all waves in flight, one per rank


def test_throughput_client_against_live_servers():
    """throughput_client drives two TesseractServers (experts interleaved across hosts) and reports latency / throughput"""
    import lah_b200 as lib
    from lah_b200.models import FeedforwardBlock
    from lah_b200.experiments.throughput import throughput_client as tc
    servers = []
    try:
        for _ in range(2):
            experts = {}
            for i in range(2):
                block = FeedforwardBlock(16)
                experts[f"expert{i}"] = lib.ExpertBackend(name=f"expert{i}", expert=block, opt=torch.optim.Adam(block.parameters()),
                                                          args_schema=(lib.BatchTensorProto(16),),
                                                          outputs_schema=lib.BatchTensorProto(16), max_batch_size=64)
            servers.append(lib.TesseractServer(None, experts, port=0, conn_handler_processes=4).run_in_background())
        args = tc.make_parser().parse_args(["-j", "3", "--hosts", *[f"127.0.0.1:{s.port}" for s in servers], "--block-type", "ffn",
                                            "--hid-dim", "16", "--batch-size", "4", "--layers-per-gpu", "2",
                                            "--batches-for-latency", "2", "--batches-for-throughput", "2", "--throughput-runs", "2",
                                            "--linspace-points", "2", "--max-ping", "0.001"])
        chain = tc.build_chain(args.hosts, args.layers_per_gpu)
        assert [(h.fn.uid, h.fn.port) for h in chain.hops] == [("expert0", servers[0].port), ("expert0", servers[1].port),
                                                                ("expert1", servers[0].port), ("expert1", servers[1].port)]
        rows = tc.run(args, printer=lambda *_: None)
        assert len(rows) == 2 and all(r["throughput"] > 0 and r["latency"] > 0 for r in rows)
        assert sum(s.runtime.samples_processed for s in servers) >= 2 * 4 * (1 + 2 + 2 * 3 * 3)
    finally:
        for s in servers:
            s.shutdown()


def test_rpc_throughput_baseline_two_processes():
    """torch.distributed.rpc baseline: rank 0 drives a chain hosted by rank 1 (CPU tensors on the wire)"""
    import os
    import socket
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    code = ("import sys; sys.path.insert(0, %r); import lah_b200; from lah_b200.experiments.throughput import rpc_throughput as r; "
            "out = r.run(r.make_parser().parse_args(sys.argv[1:])); print('RESULT', out)") % root
    common = ["--world-size", "2", "--block-type", "ffn", "--hid-dim", "16", "--batch-size", "4", "--layers-per-gpu", "3",
              "--batches-for-latency", "1", "--batches-for-throughput", "2", "--throughput-runs", "2"]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
    procs = [subprocess.Popen([sys.executable, "-c", code, "--rank", str(r), *common], env=env, stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in (0, 1)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = [l for l in outs[0].splitlines() if l.startswith("RESULT")][0]
    assert "None" not in line and float(line.split("(")[1].split(",")[0]) > 0


def test_convergence_notebooks_execute_in_smoke_mode(tmp_path, monkeypatch):
    """the three notebooks are real experiments (config -> model -> async trainers -> curve -> pickle): run their code cells"""
    import glob
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    monkeypatch.setenv("LAH_NB_SMOKE", "1")
    monkeypatch.setenv("LAH_NB_LOGDIR", str(tmp_path))
    monkeypatch.chdir(root)
    notebooks = sorted(glob.glob(os.path.join(root, "learning-at-home_b200", "experiments", "convergence", "*.ipynb")))
    assert len(notebooks) == 3
    for path in notebooks:
        nb = json.load(open(path))
        cells = ["".join(c["source"]) for c in nb["cells"] if c["cell_type"] == "code"]
        assert len(cells) >= 4
        scope = {"__file__": path}
        for src in cells:
            exec(compile(src, os.path.basename(path), "exec"), scope)
        assert len(scope["train_history"]) >= 6 and scope["val_history"] and "acc" in scope["val_history"][-1]
    assert len(list(tmp_path.iterdir())) == 3
