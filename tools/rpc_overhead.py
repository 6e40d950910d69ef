"""CPU measurement of the TCP fallback path: round trip through a near-no-op expert (compare SURVEY.md 0.3 / BASELINE.md:
reference ~8 ms floor, ~105 ms for 8 MiB each way on the same container)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

import lib


class Scale(nn.Module):   # near-no-op expert with one parameter (ExpertBackend wants an optimizer)
    def __init__(self):
        super().__init__()
        self.s = nn.Parameter(torch.ones(()))

    def forward(self, x):
        return x * self.s


def main():
    results = {}
    for name, shape in (("16 KiB", (4, 1024)), ("8 MiB", (2048, 1024))):
        expert = Scale()
        be = lib.ExpertBackend(name="e", expert=expert, opt=torch.optim.SGD(expert.parameters(), lr=0.0),
                               args_schema=(lib.BatchTensorProto(shape[1]),), outputs_schema=lib.BatchTensorProto(shape[1]),
                               max_batch_size=4096)
        srv = lib.TesseractServer(None, {"e": be}, port=0, conn_handler_processes=2, sender_threads=1)
        srv.run_in_background()
        remote = lib.RemoteExpert("e", "127.0.0.1", srv.port)
        x = torch.randn(*shape)
        with torch.no_grad():
            for _ in range(3):
                remote(x)
            t0 = time.time()
            n = 20
            for _ in range(n):
                remote(x)
            ms = (time.time() - t0) / n * 1e3
        srv.shutdown()
        nbytes = x.numel() * 4
        results[name] = dict(round_trip_ms=round(ms, 2), effective_GBps_each_way=round(nbytes / (ms / 1e3) / 1e9, 3))
        print(name, results[name], flush=True)
    return results


if __name__ == "__main__":
    main()
