"""channel_b200 - B200-native hot path of davecats/channel behind a C ABI.

Python here is the host-side mirror of the reference's driver interface (PROGRAM channel /
MODULE dnsdata) used by tests and benchmarks; the product is libchannel_b200.so.
"""
from .dnsdata import DnsIn, Channel, read_dnsin, RK1_rai, RK2_rai, RK3_rai  # noqa: F401
from ._lib import ChannelB200Error  # noqa: F401
