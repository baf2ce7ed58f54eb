"""``python -m photon_b200.node --server HOST:PORT`` — one machine of a cross-host federation (the reference's ``flower-supernode``).

Connects to the fleet link of the server process (``photon.topology=nodes`` + ``photon.fleet.n_remote_nodes``), receives the run's
config at registration, starts one worker per local GPU (DDP / ZeRO inside the node) and serves fit / evaluate / broadcast messages
until the server closes the fleet. Parameters travel through the S3 bucket when ``S3_ENDPOINT_URL`` + ``AWS_*`` are set on both
sides, inline otherwise. ``PHOTON_FLEET_TOKEN`` (both sides) authenticates the node."""
from __future__ import annotations

import argparse


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--server", required=True, help="host:port of the server's fleet link (photon.fleet.address)")
    ap.add_argument("--n-workers", type=int, default=None, help="worker processes (default: one per visible GPU, 1 on a CPU box)")
    ap.add_argument("--devices", default=None, help="comma-separated CUDA device indices for the workers (default: all visible)")
    ap.add_argument("--tls-ca", default=None, help="CA certificate: connect with TLS (env PHOTON_FLEET_TLS_CA)")
    ap.add_argument("--max-idle-s", type=float, default=None, help="leave when the server has been unreachable for this long")
    ap.add_argument("--per-gpu", action="store_true",
                    help="register ONE NODE PER GPU (each trains its own client) instead of one node whose GPUs collaborate on a client: "
                         "the right shape for models that fit a GPU — a box then trains as many clients at once as it has GPUs "
                         "(the server's photon.fleet.n_remote_nodes counts these nodes)")
    ap.add_argument("--spmd", action="store_true",
                    help="this process is ONE RANK of a box that joins as a single node running the SPMD runtime inside (fused NVLink "
                         "aggregation over the box's GPUs, one pre-aggregated model per round towards the server): launch one per GPU "
                         "with torchrun or `python -m photon_b200.launch --nproc N -m photon_b200.node -- --server host:port --spmd`")
    a = ap.parse_args()
    if a.spmd:
        from photon_b200.server.box_node import serve_box

        serve_box(a.server, tls_ca=a.tls_ca, max_idle_s=a.max_idle_s)
        return
    from photon_b200.server.grpc_fleet import serve_node
    from photon_b200.utils.core import get_n_cuda_devices

    devices = [int(x) for x in a.devices.split(",")] if a.devices else (list(range(get_n_cuda_devices())) or None)
    if a.per_gpu and devices and len(devices) > 1:
        import multiprocessing as mp

        ctx = mp.get_context("spawn")
        procs = [ctx.Process(target=serve_node, args=(a.server,), kwargs=dict(n_workers=1, devices=[d], tls_ca=a.tls_ca, max_idle_s=a.max_idle_s),
                             name=f"photon-node-gpu{d}") for d in devices]
        for p in procs:
            p.start()
        for p in procs:
            p.join()
        raise SystemExit(max((p.exitcode or 0) for p in procs))
    serve_node(a.server, n_workers=a.n_workers or (len(devices) if devices else 1), devices=devices, tls_ca=a.tls_ca, max_idle_s=a.max_idle_s)


if __name__ == "__main__":
    main()
