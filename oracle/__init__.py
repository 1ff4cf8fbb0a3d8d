"""oracle/ — CPU restatements used ONLY as the checker (tests/, smoke(), bench.py's
cpu_baseline and --impl reference legs).  The product package never imports this."""
