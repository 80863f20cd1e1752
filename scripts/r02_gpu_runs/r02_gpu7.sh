#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 900 python scripts/probe_lex_neumann.py mixed 300 450 650 900 2>&1 | tail -20
timeout 600 python scripts/probe_lex_neumann.py dirichlet 450 900 2>&1 | tail -8
