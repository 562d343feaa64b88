"""Minimal pure-torch stand-in for the third-party ``fast_transformers`` package (absent from
/root/reference and from this image).  TEST INFRASTRUCTURE: lets the reference
stage2_accompaniment/model/music_performer.py import UNMODIFIED so its glue can be checked
against oracle/performer_oracle.py.  The maths is the oracle's restatement."""
