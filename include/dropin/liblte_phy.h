// Forwarder: put -I<repo>/include/dropin in front of the reference's liblte/hdr and a caller's own `#include "liblte_phy.h"` resolves to
// the B200 drop-in surface (include/liblte_phy_b200.h), which then takes the common and RRC types from the reference's own headers.
#include "../liblte_phy_b200.h"
