#!/usr/bin/env bash
# The reference's default federated smoke run: MPT-125M, 1 client, 10 rounds, 1 local step   (ref: scripts/photon_llm_125M.sh)
exec bash "$(dirname "${BASH_SOURCE[0]}")/photon_llm.sh" 125M
