#!/usr/bin/env python3
"""Batch synthesis from a metadata file - the entry point the reference's README documents
(README.md:72-92) but which its tree folded into `synthesize.py --metadata-file`.  Thin shim: same flags
(`--text-file`, `--input-dir`, `--output-dir`, `--ckpt-path`, `--cfg-path`, `--nsteps-*`, `--temp-*`,
`--device`, `--batch-size`), metadata mode only."""
import sys

import synthesize


def main():
    parser = synthesize.build_arg_parser()
    args = parser.parse_args()
    if args.metadata_file is None:
        parser.error("--text-file/--metadata-file is required")
    if args.prompt_list is not None:
        parser.error("synthesize_via_metadata.py runs metadata mode only; use synthesize.py for --prompt-list")
    return synthesize.main(args)


if __name__ == "__main__":
    sys.exit(0 if main() is not None else 1)
