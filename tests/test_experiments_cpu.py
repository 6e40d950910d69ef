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
