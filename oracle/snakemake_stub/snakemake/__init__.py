"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for the `snakemake` package so that the
reference's scripts/calculate-optimal-parameters.py can be run stand-alone (it only needs
snakemake.shell.shell.{executable,prefix} and snakemake.io.Namedlist through
tailseeker/powersnake.py:31,64)."""
