"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/mppi_oracle.cpp). Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by the product package."""
from .binding import *  # noqa: F401,F403
