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
    a = ap.parse_args()
    from photon_b200.server.grpc_fleet import serve_node
    from photon_b200.utils.core import get_n_cuda_devices

    devices = [int(x) for x in a.devices.split(",")] if a.devices else (list(range(get_n_cuda_devices())) or None)
    serve_node(a.server, n_workers=a.n_workers or (len(devices) if devices else 1), devices=devices, tls_ca=a.tls_ca, max_idle_s=a.max_idle_s)


if __name__ == "__main__":
    main()
