// empty stand-in (oracle only): included by the reference header, unused by the extraction path
#pragma once
