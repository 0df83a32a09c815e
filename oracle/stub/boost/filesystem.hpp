// STUB (oracle/stub): nothing of boost::filesystem / boost::function is used on the compiled paths
#pragma once
