"""Network helpers for launchers: local IP / interface / free port discovery.

Capability parity with the reference's ``src/network_utils.h:28-255`` (``GetIP``, ``GetAvailableInterfaceAndIP``,
``GetAvailablePort``). The data path never uses sockets (peers share memory over NVLink); these exist for
rendezvous (``MASTER_ADDR`` / ``MASTER_PORT`` of ``torch.distributed``) and for the launcher.
"""
from __future__ import annotations

import array
import fcntl
import socket
import struct
from typing import List, Optional, Tuple


def list_interfaces() -> List[Tuple[str, str]]:
    """(interface, IPv4) pairs of this host, loopback last."""
    out: List[Tuple[str, str]] = []
    try:
        s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        nbytes = 128 * 40
        names = array.array("B", b"\0" * nbytes)
        ptr = names.buffer_info()[0]
        outbytes = struct.unpack("iL", fcntl.ioctl(s.fileno(), 0x8912, struct.pack("iL", nbytes, ptr)))[0]  # SIOCGIFCONF
        raw = names.tobytes()
        for i in range(0, outbytes, 40):
            name = raw[i:i + 16].split(b"\0", 1)[0].decode()
            ip = socket.inet_ntoa(raw[i + 20:i + 24])
            out.append((name, ip))
        s.close()
    except OSError:
        pass
    if not out:
        out = [("lo", "127.0.0.1")]
    out.sort(key=lambda p: p[1].startswith("127."))
    return out


def get_ip(interface: str) -> Optional[str]:
    """IPv4 address of a named interface (None when it does not exist)."""
    for name, ip in list_interfaces():
        if name == interface:
            return ip
    return None


def get_available_interface_and_ip() -> Tuple[str, str]:
    """First non-loopback interface with an IPv4 address; falls back to loopback (single-node boxes)."""
    return list_interfaces()[0]


def get_available_port(host: str = "127.0.0.1") -> int:
    """A TCP port that was free at the time of the call."""
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        s.bind((host, 0))
        return s.getsockname()[1]
