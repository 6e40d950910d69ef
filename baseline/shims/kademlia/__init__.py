"""Import-only shim (harness) for the `kademlia` package (not installable offline); the throughput experiment of the
reference runs with network=None and never touches the DHT."""
